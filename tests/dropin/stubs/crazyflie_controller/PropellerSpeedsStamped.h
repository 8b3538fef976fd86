#pragma once
#include <std_msgs/String.h>
namespace crazyflie_controller {   // msg/PropellerSpeedsStamped.msg: int32 fields -- assigning a double truncates
struct PropellerSpeedsStamped { std_msgs::Header header; int32_t w1 = 0, w2 = 0, w3 = 0, w4 = 0; };
}
