// gwz_blocking.cuh -- error bars of a sampled series: resample! (src/kinetic-vqmc/state.jl:194-231) followed by
// Rimu.StatsTools.blocking_analysis (Flyvbjerg-Petersen pair averaging + the M test of Jonsson, PRE 98, 043304;
// called at state.jl:53,273).  ONE implementation, compiled for the device (series_blocking_kernel, gwz_blocking.cu:
// one walker per thread, series read from HBM) and for the host (gwz_resample / gwz_blocking_analysis on
// caller-supplied vectors, gwz_stats.cpp).
//
// Formulation (differs from the reference's, same numbers):
//  * resample_stream walks the OUTPUT slots i = 1..len; slot i is either a whole slot covered by the interval that is
//    still spilling over (reference: first inner loop) or is completed by consuming input intervals (second inner
//    loop).  Every floating-point operation of state.jl:205-228 is kept, in the same order, without contraction, so
//    the resampled series is bit-identical to the reference algorithm's for the same total time (the host entry
//    point uses this form).  resample_by_input is the same computation driven by the input intervals: the device
//    kernel uses it so that all walkers read the same step of their series at the same time.
//  * on the device the resampled values go through a scratch buffer once (phase A writes them, phase B reads them
//    in lockstep); each is pushed into a BlockingStream, a binary counter of pair averages
//    that keeps five centred moments per level (sum d, sum d^2, sum d_t d_{t+1}, first d, last d; d = v - centre,
//    centre = the time-weighted mean).  From them finalize() gets, per level, exactly what Rimu's blocks_with_m
//    computes: n, var(v; corrected), autocovariance(v, 1; corrected), std_err, std_err_err, the M_j statistic; then
//    the M test as in the listing published with the paper (quantile indexed by the level: M_k < q[k]).
// PARITY UNPINNED for blocking_analysis (Rimu's source is not available here); checked against the independent
// numpy checker tests/blocking_checker.py written from the published listing.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define GWZ_HD __host__ __device__ __forceinline__
#else
#define GWZ_HD inline
#endif

namespace gwz {

constexpr int BLK_FIELDS = 5;       // s1, s2, sp, first, last
constexpr int BLK_MAX_LEVELS = 40;  // series of up to 2^40 values

// uncontracted FP64 arithmetic (the reference is plain Julia: no FMA contraction)
GWZ_HD double bl_mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
GWZ_HD double bl_add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
GWZ_HD double bl_sub(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
GWZ_HD double bl_div(double a, double b) {
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}

// a / b for a loop-invariant b with rb = 1 / b (correctly rounded): one Newton correction of a * rb.  By Markstein's
// theorem the result is the correctly rounded quotient (the exceptional divisors, mantissa all ones, give at most one
// ulp), i.e. the value the reference's division produces -- without the ~30-instruction IEEE division sequence.
GWZ_HD double bl_div_by(double a, double b, double rb) {
#ifdef __CUDA_ARCH__
    const double q = __dmul_rn(a, rb);
    const double e = __fma_rn(-b, q, a);
    return __fma_rn(e, rb, q);
#else
    (void)rb;
    return a / b;
#endif
}

// resample!(result, residence_times, values) with length(result) = len, streamed: emit(v) is called len times, in
// slot order.  load(k, tau, val) fetches interval k (0-based, k < n); `total` = sum(residence_times).
template <class LOAD, class EMIT>
GWZ_HD void resample_stream(uint64_t n, uint64_t len, double total, LOAD&& load, EMIT&& emit) {
    double t = 0.0;    // state.jl:204  end of the consumed intervals, in slot units
    double mu = 0.0;   // :206  integral over the part of slot i seen so far
    double val = 0.0;  // :207  value of the interval that is spilling over
    uint64_t k = 0;
    const double dlen = (double)len;
    const double rtotal = bl_div(1.0, total);
    for (uint64_t i = 1; i <= len; ++i) {
        const double di = (double)i;
        if (t >= di) {  // :212-216  a whole slot inside the spilling interval
            emit(val);
            mu = bl_sub(mu, val);
        } else {
            while (t < di && k < n) {  // :218-224  consume intervals until slot i is full
                double tau;
                load(k, tau, val);
                const double step = bl_div_by(bl_mul(dlen, tau), total, rtotal);  // :220  len * tau / total
                const double room = bl_sub(di, t);
                mu = bl_add(mu, bl_mul(val, step < room ? step : room));  // :221
                t = bl_add(t, step);
                ++k;
            }
            emit(mu);                        // :227
            mu = bl_mul(bl_sub(t, di), val);  // :228  carry the overshoot into the next slot
        }
    }
}

// The same resampling driven by the INPUT intervals k = 0..n-1 (the loop nesting of state.jl:209-230 turned inside
// out; every floating-point operation and its order per series are unchanged, so the values are identical to
// resample_stream's).  On the device all threads then read interval k of their walkers at the same time (coalesced,
// prefetchable); what varies between threads is how many slots an interval completes, i.e. how often emit() runs.
// emit(i, v): slot i (0-based) of the result is v.
template <class LOAD, class EMIT>
GWZ_HD void resample_by_input(uint64_t n, uint64_t len, double total, LOAD&& load, EMIT&& emit) {
    double t = 0.0, mu = 0.0, val = 0.0;
    uint64_t i = 1;  // the slot being filled (1-based like the reference)
    const double dlen = (double)len;
    const double rtotal = bl_div(1.0, total);
    for (uint64_t k = 0; k < n && i <= len; ++k) {
        double tau;
        load(k, tau, val);
        const double step = bl_div_by(bl_mul(dlen, tau), total, rtotal);  // state.jl:220
        const double room = bl_sub((double)i, t);
        mu = bl_add(mu, bl_mul(val, step < room ? step : room));  // :221
        t = bl_add(t, step);
        if (t >= (double)i || k + 1 == n) {  // slot i is full, or the intervals are exhausted (:218 exits)
            emit(i - 1, mu);                               // :227
            mu = bl_mul(bl_sub(t, (double)i), val);        // :228
            ++i;
            while (t >= (double)i && i <= len) {  // :212-216 whole slots inside the interval that spills over
                emit(i - 1, val);
                ++i;
                mu = bl_sub(mu, val);
            }
        }
    }
    while (i <= len) {  // only reached through rounding of t: the last slots take what is left (:227-229)
        emit(i - 1, mu);
        mu = bl_mul(bl_sub(t, (double)i), val);
        ++i;
    }
}

struct BlockingOut {
    double mean, err, err_err, p_cov;
    int k;           // 1-based level accepted by the M test; 0: single value; -1: "M test failed, more data needed"
    int64_t blocks;  // length of the accepted level
};

// chi-square 0.99 quantiles for 1..30 degrees of freedom, the table printed in Jonsson's listing
#define GWZ_CHI2_Q99_TABLE                                                                                          \
    {6.634897,  9.210340,  11.344867, 13.276704, 15.086272, 16.811894, 18.475307, 20.090235, 21.665994, 23.209251, \
     24.724970, 26.216967, 27.688250, 29.141238, 30.577914, 31.999927, 33.408664, 34.805306, 36.190869, 37.566235, \
     38.932173, 40.289360, 41.638398, 42.979820, 44.314105, 45.641683, 46.962942, 48.278236, 49.587884, 50.892181}

GWZ_HD int blocking_levels(uint64_t len) {
    int l = 0;
    while (l + 1 < 63 && (1ull << (l + 1)) <= len) ++l;  // floor(log2(len))
    return l;
}

// STORE: double& operator()(int level, int field); levels x BLK_FIELDS doubles, zero-initialised by the caller.
template <class STORE>
struct BlockingStream {
    STORE st;
    int levels;
    uint64_t pushed;
    double centre;

    double a0, a1, a2, a3, a4;  // the moments of level 0 (touched by every push) stay in registers

    GWZ_HD BlockingStream(STORE s, int lv, double c)
        : st(s), levels(lv), pushed(0), centre(c), a0(0.0), a1(0.0), a2(0.0), a3(0.0), a4(0.0) {}

    GWZ_HD void push(double v) {
        double d = v - centre;
        uint64_t cnt = pushed++;
        {  // level 0
            const double last = a4;
            a0 += d;
            a1 += d * d;
            if (cnt == 0) a3 = d;
            else a2 += last * d;
            a4 = d;
            if (!(cnt & 1ull)) return;  // the first value of a pair waits for its partner
            d = 0.5 * (last + d);       // Rimu `blocker`: 1/2 (v[2i-1] + v[2i]); an odd last value is dropped
            cnt >>= 1;
        }
        for (int l = 1; l < levels; ++l) {  // cnt = values already seen by level l
            const double last = st(l, 4);
            st(l, 0) += d;
            st(l, 1) += d * d;
            if (cnt == 0) st(l, 3) = d;
            else st(l, 2) += last * d;
            st(l, 4) = d;
            if (!(cnt & 1ull)) break;
            d = 0.5 * (last + d);
            cnt >>= 1;
        }
    }

    // level-l moments update with one value d (cnt = values already seen by that level): registers or store
    static GWZ_HD void upd(double& m0, double& m1, double& m2, double& m3, double& m4, double d, bool first) {
        const double last = m4;
        m0 += d;
        m1 += d * d;
        if (first) m3 = d;
        else m2 += last * d;
        m4 = d;
    }

    // Four consecutive values at once, for a stream whose push count is a multiple of four (the device's phase B,
    // where all threads are at the same value index): the carry pattern of levels 0 and 1 is then static -- level 0
    // sees all four, level 1 their two pair averages (b0..b4 stay in registers next to a0..a4), level 2 the average
    // of the four -- and only levels >= 2 touch the store, once per four values.  Needs levels >= 2.
    GWZ_HD void push4(double v0, double v1, double v2, double v3, double& b0, double& b1, double& b2, double& b3,
                      double& b4) {
        const uint64_t cnt = pushed;  // multiple of 4
        pushed += 4;
        const double d0 = v0 - centre, d1 = v1 - centre, d2 = v2 - centre, d3 = v3 - centre;
        upd(a0, a1, a2, a3, a4, d0, cnt == 0);
        upd(a0, a1, a2, a3, a4, d1, false);
        upd(a0, a1, a2, a3, a4, d2, false);
        upd(a0, a1, a2, a3, a4, d3, false);
        const double p0 = 0.5 * (d0 + d1), p1 = 0.5 * (d2 + d3);
        upd(b0, b1, b2, b3, b4, p0, cnt == 0);
        upd(b0, b1, b2, b3, b4, p1, false);
        double d = 0.5 * (p0 + p1);
        uint64_t c = cnt >> 2;  // values already seen by level 2
        for (int l = 2; l < levels; ++l) {
            const double last = st(l, 4);
            st(l, 0) += d;
            st(l, 1) += d * d;
            if (c == 0) st(l, 3) = d;
            else st(l, 2) += last * d;
            st(l, 4) = d;
            if (!(c & 1ull)) break;
            d = 0.5 * (last + d);
            c >>= 1;
        }
    }
    GWZ_HD void flush_level1(double b0, double b1, double b2, double b3, double b4) {
        if (levels > 1) {
            st(1, 0) = b0;
            st(1, 1) = b1;
            st(1, 2) = b2;
            st(1, 3) = b3;
            st(1, 4) = b4;
        }
    }

    // len = number of values pushed (>= 2); q99 = the quantile table above
    GWZ_HD void finalize(uint64_t len, const double* q99, BlockingOut& o) {
        if (levels > 0) {
            st(0, 0) = a0;
            st(0, 1) = a1;
            st(0, 2) = a2;
            st(0, 3) = a3;
            st(0, 4) = a4;
        }
        o.mean = centre + a0 / (double)len;
        o.err = o.err_err = o.p_cov = NAN;
        o.k = -1;
        o.blocks = 0;
        for (int l = 0; l < levels; ++l) {  // Rimu blocks_with_m, level l+1
            const double n = (double)(len >> l);
            const double s1 = st(l, 0), s2 = st(l, 1), sp = st(l, 2), first = st(l, 3), last = st(l, 4);
            const double m = s1 / n;  // mean of the level, relative to the centre
            double ss = s2 - n * m * m;
            ss = ss > 0.0 ? ss : 0.0;
            const double cs = sp - m * (2.0 * s1 - first - last) + (n - 1.0) * m * m;
            const double var = ss / (n - 1.0);     // var(v; corrected = true)
            const double gamma = cs / (n - 1.0);   // autocovariance(v, 1; corrected = true)
            const double se = sqrt(var / n);
            const double x = (n - 1.0) * var / (n * n) + gamma;
            st(l, 0) = n * x * x / (var * var);    // M_j; NaN for a constant series, like the reference
            st(l, 1) = se;
            st(l, 2) = se / sqrt(2.0 * (n - 1.0));
        }
        double tail = 0.0;
        for (int l = levels - 1; l >= 0; --l) {  // M_k = sum_{j >= k} M_j  (reverse(cumsum(reverse(mj))))
            tail += st(l, 0);
            st(l, 0) = tail;
        }
        for (int l = 0; l + 1 < levels; ++l) {  // mtest: k = 1 .. length(m) - 1, m[k] < q[k]
            if (st(l, 0) < q99[l < 30 ? l : 29]) {
                o.k = l + 1;
                o.err = st(l, 1);
                o.err_err = st(l, 2);
                o.p_cov = o.err * o.err;
                o.blocks = (int64_t)(len >> l);
                break;
            }
        }
    }
};

}  // namespace gwz
