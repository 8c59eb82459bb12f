// A C++ host stepping many robots per tick through host/FormCPipeline.hpp (plain g++, no CUDA headers):
// what a batch version of AMR_code_DART/Controller.cpp:105-106,346-348 looks like.  Prints one line per tick with the
// CoM of robot 0; exits 2 with the library's message if no B200 is present (there is no CPU fallback).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../quadruped_gait_generation_ismpc_b200/host/FormCPipeline.hpp"

int main(int argc, char** argv)
{
    const int n = 64, depth = 2, ticks = argc > 1 ? std::atoi(argv[1]) : 4, n_steps = 40;
    ismpc_formc_model_t model{};
    model.dt = 0.01; model.dtc = 0.01; model.mass = 50.0; model.g = 9.81;                 // parameters.cpp:9-45
    model.q_p = 1005000.0; model.q_v = 100.0; model.q_u = 0.01; model.fz_max = 1e4; model.N = 100;
    std::vector<double> plan((size_t)n_steps * 4);                                         // Controller.cpp:89-97
    for (int i = 0; i < n_steps; ++i) {
        plan[4 * i + 0] = (i - 1) * 0.2; plan[4 * i + 1] = (i % 2 ? -0.08 : 0.08); plan[4 * i + 2] = 0.0; plan[4 * i + 3] = 45.0 * i;
    }
    try {
        ismpc_host::FormCPipeline p(0, n, depth, model, 35, 10, plan.data(), n_steps);
        for (int k = 0; k < ticks + depth; ++k) {
            const int s = p.acquire();
            if (k >= depth) std::printf("%d %.9f %.9f %.9f\n", k - depth, p.out(s)[0].next.com_pos[0], p.out(s)[0].next.com_pos[1], p.out(s)[0].next.com_pos[2]);
            if (k >= ticks) continue;
            for (int i = 0; i < n; ++i) {
                ismpc_state_t& st = p.state(s)[i]; ismpc_walk_t& wk = p.walk(s)[i]; ismpc_formc_inst_t& in = p.inst(s)[i];
                st = ismpc_state_t{}; st.com_pos[2] = 0.69;
                wk = ismpc_walk_t{}; wk.sim_time = k; wk.mpc_iter = k % 45; wk.control_iter = k % 45; wk.footstep_counter = 2;
                in = ismpc_formc_inst_t{}; in.com_height = 0.69; in.box_w = 0.09; in.box_w_init = 2.0; in.S = 35; in.F_ds = 10;
                in.plan_first_row = 0; in.n_steps = n_steps;
            }
            p.submit(s);
        }
        p.wait_all();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 2;
    }
    return 0;
}
