#pragma once
#include <geometry_msgs/Twist.h>
namespace crazyflie_controller {   // msg/CrazyflieState.msg
struct CrazyflieState { geometry_msgs::Vector3 pos; geometry_msgs::Quaternion quat; geometry_msgs::Vector3 vel, rates; };
}
