#pragma once
#include <boost/thread.hpp>
#include <cstdint>
namespace dynamic_reconfigure {
template <class C> struct Server
{
    typedef boost::function<void(C &, uint32_t)> CallbackType;
    void setCallback(const CallbackType &f)
    {   // roscpp invokes the callback once with the default configuration and level ~0
        C c;
        f(c, ~0u);
    }
};
}
