#pragma once
#include <geometry_msgs/Twist.h>
namespace geometry_msgs {
struct PointStamped { std_msgs::Header header; Point point; };
}
