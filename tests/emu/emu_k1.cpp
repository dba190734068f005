// TEST INFRASTRUCTURE — runs the K1 reach-set kernel on the CPU through the fiber emulator
// (tests/emu/cuda_emu.h) so its logic can be checked against the oracle without a GPU.
// Built by tests/test_k1_emu.py with g++ -O1 -ffp-contract=off; never part of the product.
#include "cuda_emu.h"

#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../armour_b200/csrc/k1_reachsets.cuh"

using namespace armour;

extern "C" int emu_k1_build(int model_id, int T, double thr, const double* k_range, double mass_unc, double inertia_unc,
                            const double* q0, const double* qd0, const double* qdd0, const int* units, int nunits,
                            int capL, int capU, int arena_words, int tab_s_bytes, int* link_n, double* link_c,
                            unsigned short* link_key, double* link_g, int* u_n, double* u_c, double* u_r,
                            unsigned short* u_key, double* u_g, double* torque_radius, double* link_gens, int* stats) {
    RobotConstants rc = make_robot_constants(model_id);
    rc.num_time_steps = T;
    rc.simplify_threshold = thr;
    for (int i = 0; i < NF; i++) rc.k_range[i] = k_range[i];
    if (mass_unc >= 0) rc.mass_uncertainty = mass_unc;
    if (inertia_unc >= 0) rc.inertia_uncertainty = inertia_unc;
    c_robot = rc;
    Batch B;
    std::memset(&B, 0, sizeof(B));
    B.nprob = 1;
    B.epoch = 0;
    B.hp_slow = nullptr;
    B.T = T;
    B.NJ = rc.num_joints;
    B.O = 0;
    B.capL = capL;
    B.capU = capU;
    B.q0 = q0;
    B.qd0 = qd0;
    B.qdd0 = qdd0;
    B.link_n = link_n;
    B.link_c = link_c;
    B.link_key = link_key;
    B.link_g = link_g;
    B.u_n = u_n;
    B.u_c = u_c;
    B.u_r = u_r;
    B.u_key = u_key;
    B.u_g = u_g;
    B.torque_radius = torque_radius;
    B.link_gens = link_gens;
    std::vector<double> link_r(size_t(T) * MAXJ * 3);
    B.link_r = link_r.data();
    int status = 0, work = 0;
    B.status = &status;
    k1::K1Params P;
    P.B = B;
    P.work = &work;
    const int emu_capw = std::getenv("EMU_CAPW") ? std::atoi(std::getenv("EMU_CAPW")) : 4096;  // developer knob
    const int gscr_words = 2 * MAXJ * (9 + emu_capw * 4);
    std::vector<double> gscr(gscr_words);
    const int gtab_bytes = std::getenv("EMU_GTAB") ? std::atoi(std::getenv("EMU_GTAB")) : (1 << 22);
    std::vector<char> gtab(gtab_bytes, 0);
    P.gscr = gscr.data();
    P.gscr_words = gscr_words;
    P.fn_words = gscr_words / 2;
    P.gtab = gtab.data();
    P.gtab_bytes = gtab_bytes;
    P.arena_words = arena_words;
    P.tab_s_bytes = tab_s_bytes;
    P.units = units;
    P.nunits = nunits;
    P.stats = stats;
    P.mbox = nullptr;
    P.mbox_words = 0;
    P.group_bytes = k1::K1_FIXED_BYTES + arena_words * 8 + tab_s_bytes;
    const size_t smem = 16 + size_t(P.group_bytes);
    emu::launch(dim3(1), dim3(k1::NT), smem, [&]() { k1::k_reachsets(P); });
    for (char ch : gtab)
        if (ch) return -100;  // the global table pool must be left all-zero
    return status;
}

// claim order of the MG latency configuration (armour_b200/csrc/k1_reachsets.cuh: mg_task_list), for the CPU check
// that it is a topological order: out[i] = kind << 8 | joint, returns the number of tasks
extern "C" int emu_mg_task_list(int NJ, unsigned short* out) { return k1::mg_task_list(out, NJ); }
