#pragma once
#include <cstdint>
namespace crazyflie_controller {   // msg/PropellerSpeeds.msg
struct PropellerSpeeds { int32_t w1 = 0, w2 = 0, w3 = 0, w4 = 0; };
}
