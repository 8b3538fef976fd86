// Stand-in for <ros/ros.h>: just enough of roscpp for crazyflie_controller/src/acados_mpc.cpp to compile unmodified.
// Publishers append every published message to a per-type log that the test driver reads.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

namespace boost {
template <class T> using shared_ptr = std::shared_ptr<T>;
}
namespace ros {
struct Time
{
    double t = 0.0;
    static Time now() { return Time(); }
};
template <class M> std::vector<M> &published()
{
    static std::vector<M> log;
    return log;
}
struct Publisher
{
    template <class M> void publish(const M &m) const { published<M>().push_back(m); }
};
struct Subscriber
{
};
struct NodeHandle
{
    explicit NodeHandle(const std::string & = std::string()) {}
    template <class M> Publisher advertise(const std::string &, int) { return Publisher(); }
    template <class... A> Subscriber subscribe(const std::string &, int, A...) { return Subscriber(); }
    static std::string &param_ref_traj()
    {
        static std::string s;
        return s;
    }
    bool getParam(const std::string &name, std::string &v) const
    {
        if (name == "ref_traj") { v = param_ref_traj(); return true; }
        return false;
    }
};
inline void init(int &, char **, const std::string &) {}
inline void spin() {}
}  // namespace ros
#define ROS_INFO(...) do { } while (0)
#define ROS_WARN(...) do { } while (0)
#define ROS_DEBUG(...) do { } while (0)
#define ROS_INFO_STREAM(x) do { } while (0)
