#pragma once
#include <crazyflie_controller/CrazyflieState.h>
#include <crazyflie_controller/PropellerSpeeds.h>
#include <vector>
namespace crazyflie_controller {   // msg/CrazyflieOpenloopTraj.msg
struct CrazyflieOpenloopTraj { std_msgs::Header header; double cpu_time = 0; std::vector<CrazyflieState> states; std::vector<PropellerSpeeds> controls; };
}
