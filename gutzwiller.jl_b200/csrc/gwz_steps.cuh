// gwz_steps.cuh -- one kinetic_sample! (src/kinetic-vqmc/qmc.jl:14-44) per call, two device paths.
#pragma once
#include "gwz_device.cuh"

namespace gwz {

// ------------------------------------------------------------------------------------------
// one continuous-time step, production path (bond classes by popcount)
// ------------------------------------------------------------------------------------------
template <int NW, int KC>
__device__ __forceinline__ void step_bitparallel(uint32_t (&a)[NW], const DevModel& dm, const DevTables& tb,
                                                 const StepScratch<NW, KC>& sc, double u01, double& tau, double& eloc,
                                                 double& gratio, int& nhops, int* ref_hop_index = nullptr) {
    uint32_t e[NW];
    extend<NW, KC>(a, dm, e);
    BondMasks<NW, KC> bm;
    classify_bonds<NW, KC>(e, dm, bm);
    BondSums bs;
    double bigS;
    const UnaryBig<NW, KC> big{e, bm.big, dm, sc};
    const double reg_total = bond_sums<NW, KC, false>(tb, bm, SmemCumSink{sc.cum + threadIdx.x}, big, bs, bigS);
    stage_masks<NW, KC>(bm, sc);
    const double D = dm.u * (double)bs.d;  // diagonal_element (qmc.jl:20)
    tau = 1.0 / bs.S;                      // |psi1| / sum |psi2|   (qmc.jl:33)
    eloc = D - dm.t * bs.E;                // (diag*psi1 + sum melem*psi2)/psi1  (qmc.jl:21,28,37)
    gratio = -D;                           // grad1/val1 = -diag   (gutzwiller.jl:30, qmc.jl:38)
    nhops = 2 * bs.nocc;
    double T = u01 * bs.S;                 // rand() * last(cumsum)  (qmc.jl:53)
    if (!(T < bs.S)) T = __longlong_as_double(__double_as_longlong(bs.S) - 1);
    int z;
    bool right;
    select_move<NW, KC>(tb, sc, big, reg_total, bigS, T, z, right);
    if (z >= 0) {
        if (ref_hop_index) *ref_hop_index = reference_hop_index<NW>(e, dm, z, right);
        apply_hop<NW>(a, dm, z, right);
    }
}

// ------------------------------------------------------------------------------------------
// one step in the reference's hop order with the reference's selection rule (qmc.jl:52-61)
// ------------------------------------------------------------------------------------------
template <int NW>
__device__ __forceinline__ void step_enumerate(uint32_t (&a)[NW], const DevModel& dm, const DevTables& tb,
                                               double u01, double& tau, double& eloc, double& gratio,
                                               int& nhops, int& chosen, double* rates_out) {
    const double c = tb.c;
    double S = 0.0, E = 0.0;
    long long d = 0;
    int h = 0;
    enumerate_hops<NW>(a, dm, [&](int hop, int, bool right, int nsrc, int ndst) {
        const double rate = exp(-c * (double)(ndst - nsrc + 1));
        S += rate;                                                 // residence_time_denom (qmc.jl:27)
        E += sqrt((double)nsrc * (double)(ndst + 1)) * rate;       // -melem*psi2/psi1 / t   (qmc.jl:28)
        if (right) d += (long long)nsrc * (nsrc - 1) / 2;
        if (rates_out) rates_out[hop - 1] = rate;
        h = hop;
        return false;
    });
    const double D = dm.u * (double)d;
    tau = 1.0 / S;
    eloc = D - dm.t * E;
    gratio = -D;
    nhops = h;
    // pick_random_from_cumsum: first i with u*S < cumsum_i, clamped to the last hop (App. B.7)
    const double T = u01 * S;
    double cum = 0.0;
    int zsel = -1;
    bool rsel = true;
    int csel = 0;
    enumerate_hops<NW>(a, dm, [&](int hop, int z, bool right, int nsrc, int ndst) {
        cum += exp(-c * (double)(ndst - nsrc + 1));
        zsel = z;
        rsel = right;
        csel = hop;
        return T < cum;
    });
    chosen = csel;
    if (zsel >= 0) apply_hop<NW>(a, dm, zsel, rsel);
}


}  // namespace gwz
