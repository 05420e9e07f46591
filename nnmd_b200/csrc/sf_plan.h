// sf_plan.h -- host-side description of one species' symmetry-function table (nnmd_sf_plan).
#pragma once
#include <vector>

#include "common.cuh"

namespace nnmd {

constexpr int ANG_MAX_FPL = 4;    // angular functions per lane (template values 1, 2, 4)
constexpr int REC_STRIDE = 52;    // shared-memory record stride in reals (52 mod 32 = 20: conflict-free 16 B accesses)
constexpr int REC_CLASS = 20;     // reals per lambda-class block
constexpr int REC_TAIL = 40;      // offset of the interleaved unit vectors
constexpr int ANG_BATCH = 32;     // triangles staged per batch
constexpr int NB_FIELDS = 10;     // per-neighbour staged reals: ux uy uz d 1/d fc fc' psi psi' + edge slot (int)
constexpr int EREC_STRIDE = 36;   // edge-centric record stride (36 mod 32 = 4: conflict-free 16 B accesses)
constexpr int EREC_CLASS = 16;
constexpr int64_t EDGE_WS_HEADER = 256;  // bytes at the head of the caller's edge scratch: per-call row counter

// One launch of the angular kernel: functions that share (kind, rc, eta).
struct AngGroup {
    int kind;           // NNMD_G4 | NNMD_G5
    double rc, eta;
    int pm1;            // all lambda in {+1,-1}: logs precomputed per class in the triangle record
    int fpl;            // functions per lane
    int lanes;          // G: lanes per triangle (power of two <= 32)
    int n_slots;        // lanes * fpl table entries
    int ladder = 0;     // fpl == 2 and every lane's two zetas differ by the same step: q_b = q_a * 2^(step * L)
    double ladder_step = 0;
    int edge_path = 0;  // fp32, lambda = +-1, 4 functions per lane on 16 lanes: eligible for the edge-centric kernel
    int edge_base = 0;  // first slot of this group on the function axis of the edge scratch
    // device tables in the working dtype, n_slots entries each (slot = lane*fpl + i); col = -1 marks padding
    void *d_zeta = nullptr, *d_lam = nullptr, *d_coef = nullptr;
    int *d_col = nullptr;
    std::vector<int> cols;  // host copy (tests / debugging)
    // multi-species plans: symmetric pair-weight table pw[s_j * n_species + s_k] of this group (working dtype), else NULL
    void *d_pw = nullptr;
};

struct RadialTab {
    int n = 0;          // number of G1/G2 functions
    int uniform_rc = 0; // every radial function has the same cutoff: fc is staged once per neighbour
    double rc_max = 0;
    void *d_rc = nullptr, *d_eta = nullptr, *d_rs = nullptr;  // working dtype; G1 rows carry eta = 0, rs = 0
    int *d_col = nullptr;
    void *d_wr = nullptr;  // multi-species plans: neighbour-species weight vector wr[s_j] of this group (working dtype), else NULL
};

}  // namespace nnmd

struct nnmd_sf_plan {
    int dtype;
    int nf;
    double rc_max;
    nnmd::RadialTab rad;
    std::vector<nnmd::RadialTab> ms_rads;  // multi-species plans: one radial launch per distinct neighbour-weight vector
    int n_species = 0;                     // > 0: multi-species plan (posw.w carries the species index of every atom)
    std::vector<nnmd::AngGroup> groups;
    int edge_slots = 0;  // function-axis length of the edge scratch (sum of n_slots over edge-path groups)
};
