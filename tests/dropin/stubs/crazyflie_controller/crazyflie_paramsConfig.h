#pragma once
namespace crazyflie_controller {   // cfg/crazyflie_params.cfg: the fields the node reads
struct crazyflie_paramsConfig
{
    bool enable_traj_tracking = false, enable_regulation = false;
    double xq_des = 0, yq_des = 0, zq_des = 0;
    double Wdiag_xq = 0, Wdiag_yq = 0, Wdiag_zq = 0, Wdiag_qw = 0, Wdiag_qx = 0, Wdiag_qy = 0, Wdiag_qz = 0, Wdiag_vbx = 0, Wdiag_vby = 0,
           Wdiag_vbz = 0, Wdiag_wx = 0, Wdiag_wy = 0, Wdiag_wz = 0, Wdiag_w1 = 0, Wdiag_w2 = 0, Wdiag_w3 = 0, Wdiag_w4 = 0;
};
}
