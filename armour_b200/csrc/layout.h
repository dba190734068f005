// Device-side data layout of a batch of planning problems (HBM, structure of arrays).
//
// All arrays are problem-major; inside a problem the time interval is the slow index because every
// kernel assigns a CTA to one (problem, time interval) or one (problem, block of TB intervals), so
// each CTA streams one contiguous chunk.
#pragma once
#include <cstdint>

#include "robot_constants.h"

namespace armour {

#ifndef ARMOUR_TB
#define ARMOUR_TB 8
#endif
constexpr int TB = ARMOUR_TB;      // time intervals per CTA ("chunk") in the constraint kernels
constexpr int HP_CAP = 32;         // most candidate half-spaces stored for one (link, interval, obstacle) row
constexpr int HP_OVERFLOW = 255;   // row count marker: no stored list, evaluate the row from the generators
constexpr double K_DOMAIN = 1.0 + 1e-6;        // the candidate lists are exact for |k_j| <= K_DOMAIN
constexpr double HP_RHO_SCALE = 1.0 + 3e-5;    // >= K_DOMAIN^21 (largest total degree of a link monomial)

struct Batch {
    int nprob;     // problems in flight
    int T;         // time intervals
    int NJ;        // links
    int O;         // obstacles per problem (same for the whole batch)
    int capL;      // stored k-only monomials per link reach set
    int capU;      // stored k-only monomials per torque reach set
    int epoch;     // ordinal of the build that filled the tables (> 0); status[] and the unit flags are stamped with it
    // inputs
    const double* q0;         // [p][NF]
    const double* qd0;        // [p][NF]
    const double* qdd0;       // [p][NF]
    const double* obstacles;  // [p][O][12]
    // ARMTD comparison planner only (nullptr otherwise): cos / sin models of the joint reachable set handed in instead of
    // derived from the Bezier trajectory, [p][t][NF][6] = cos centre, k-coefficient, error coefficient, then the same for sin
    const double* jrs_ext;
    // reach sets written by the build kernel (or armour_import_reachsets)
    int* link_n;              // [p][t][NJ]
    double* link_c;           // [p][t][NJ][3]
    uint16_t* link_key;       // [p][t][NJ][capL]   2 bits per k_j (reference degree hash, k-only part)
    double* link_g;           // [p][t][NJ][capL][3]
    int* u_n;                 // [p][t][NF]
    double* u_c;              // [p][t][NF]
    double* u_r;              // [p][t][NF]         radius of the reduced nominal torque PZ
    uint16_t* u_key;          // [p][t][NF][capU]
    double* u_g;              // [p][t][NF][capU]
    double* torque_radius;    // [p][j*T + t]
    double* link_gens;        // [p][t][NJ][18]     column-major 3x6
    double* link_r;           // [p][t][NJ][3]      diag of the radius block of link_gens (all an evaluation needs)
    // collision half-space candidates: records [p][t/TB][candidate][l][t%TB][o][4] = (s*Cx, s*Cy, s*Cz, b): candidate-
    // major inside a chunk, so that the q-th records of 32 consecutive rows are one contiguous KB; the number of candidates
    // per row [p][t/TB][l][t%TB][o] (HP_OVERFLOW: no list); and per problem the number of rows without a list
    double* hp_cand;
    unsigned char* hp_cnt;
    int* hp_slow;
    // outputs of the last evaluation
    double* link_sliced;      // [t][NJ][3] of problem 0 (armtd_NLP::link_sliced_center)
    int* status;              // [p] epoch * 8 + failure code of a build that overflowed a table (see failed()): evaluations
                              // of such a problem return fail-safe rows
    // optional indirection for launches over a subset of the problems (the batched solver's still-running list):
    // CTA row y works on problem plist[y]; nullptr = identity
    const int* plist;

    __host__ __device__ size_t hp_chunk() const { return size_t(HP_CAP) * 4 * NJ * TB * O; }
    __host__ __device__ int m() const { return NF * T + NJ * T * O + 4 * NF; }
    // failure code of problem p in the CURRENT build (0: none); stamps of older builds do not count
    __device__ int failed(int p) const { return (status[p] >> 3) == epoch ? (status[p] & 7) : 0; }
};

}  // namespace armour
