#pragma once
#include <geometry_msgs/Twist.h>
namespace crazyflie_controller {   // msg/CrazyflieStateStamped.msg
struct CrazyflieStateStamped { std_msgs::Header header; geometry_msgs::Vector3 pos; geometry_msgs::Quaternion quat; geometry_msgs::Vector3 vel, rates; };
typedef boost::shared_ptr<CrazyflieStateStamped> CrazyflieStateStampedPtr;
}
