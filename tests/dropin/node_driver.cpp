// Drives the reference's ROS node class without ROS.  The node source itself is NOT in this repository: it is compiled
// from where it lies, unmodified, through the computed include below (tests/dropin/build_node.py passes
// -DCF_NODE_SOURCE="\"/root/reference/crazyflie_controller/src/acados_mpc.cpp\"" and -I tests/dropin/stubs for the
// ROS / boost / Eigen headers).  Linked once against the reference's own acados build (node_ref: mints the golden
// vectors of tests/golden/node_loop_golden.npz) and once against include/ + libcfnmpc.so (node_ours: the drop-in claim
// of INTEGRATION.md section A, run on the GPU box by tests/test_gpu_node.py).
//
// usage: node_<backend> scenario.bin trajectory.txt out.bin [solver_log.bin]
//   scenario.bin  doubles: n_ticks, then per tick [cmd, xq_des, yq_des, zq_des, state(13)]
//                 cmd: -1 none, 0 dynamic-reconfigure "regulation" with the set-point, 1 "trajectory tracking"
//   out.bin       doubles per tick: published /crazyflie/acados_motvel w1..w4 (int32 fields) and /crazyflie/cmd_vel
//                 linear.x, linear.y, linear.z, angular.z
#define main cf_reference_node_main
#include CF_NODE_SOURCE
#undef main
#undef N
#undef NX
#undef NU
#undef NY
#undef NYN
#undef pi
#undef g0

#include <cstdio>

extern "C" void cf_glue_set_log(const char *path) __attribute__((weak));

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s scenario.bin trajectory.txt out.bin [solver_log.bin]\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    double nt = 0;
    if (fread(&nt, 8, 1, f) != 1) return 2;
    const int n_ticks = (int) nt;
    std::vector<double> sc((size_t) n_ticks * 17);
    if (fread(sc.data(), 8, sc.size(), f) != sc.size()) { fprintf(stderr, "short scenario file\n"); return 2; }
    fclose(f);
    if (argc > 4 && cf_glue_set_log) cf_glue_set_log(argv[4]);

    ros::NodeHandle::param_ref_traj() = argv[2];
    ros::NodeHandle n("~");
    std::string ref_traj;
    n.getParam("ref_traj", ref_traj);
    NMPC nmpc(n, ref_traj);

    FILE *out = fopen(argv[3], "wb");
    if (!out) { perror(argv[3]); return 2; }
    for (int t = 0; t < n_ticks; t++) {
        const double *r = sc.data() + (size_t) t * 17;
        if (r[0] >= 0) {
            crazyflie_controller::crazyflie_paramsConfig c;
            c.enable_regulation = r[0] == 0;
            c.enable_traj_tracking = r[0] == 1;
            c.xq_des = r[1]; c.yq_des = r[2]; c.zq_des = r[3];
            nmpc.callback_dynamic_reconfigure(c, 1);
        }
        crazyflie_controller::CrazyflieStateStampedPtr msg(new crazyflie_controller::CrazyflieStateStamped());
        const double *x = r + 4;
        msg->pos.x = x[0]; msg->pos.y = x[1]; msg->pos.z = x[2];
        msg->quat.w = x[3]; msg->quat.x = x[4]; msg->quat.y = x[5]; msg->quat.z = x[6];
        msg->vel.x = x[7]; msg->vel.y = x[8]; msg->vel.z = x[9];
        msg->rates.x = x[10]; msg->rates.y = x[11]; msg->rates.z = x[12];
        nmpc.iteration(msg);
        auto &mot = ros::published<crazyflie_controller::PropellerSpeedsStamped>();
        auto &tw = ros::published<geometry_msgs::Twist>();
        if ((int) mot.size() != t + 1 || (int) tw.size() != t + 1) { fprintf(stderr, "tick %d: nothing was published\n", t); return 3; }
        const double o[8] = {(double) mot.back().w1, (double) mot.back().w2, (double) mot.back().w3, (double) mot.back().w4,
                             tw.back().linear.x, tw.back().linear.y, tw.back().linear.z, tw.back().angular.z};
        fwrite(o, 8, 8, out);
    }
    fclose(out);
    nmpc.nmpcReset();
    return 0;
}
