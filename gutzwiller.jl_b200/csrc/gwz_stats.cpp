// gwz_stats.cpp -- resample! (src/kinetic-vqmc/state.jl:194-231) and Rimu.StatsTools.blocking_analysis for
// caller-supplied host vectors.  The algorithm is the one in gwz_blocking.cuh, shared with the device kernel that
// analyses the series recorded by the walker kernels (gwz_blocking.cu); see there for the formulation and for what
// is and is not pinned.
#include "gutzwiller_b200.h"

#include <cmath>
#include <limits>
#include <vector>

#include "gwz_blocking.cuh"

namespace {
const double kChi2Q99[30] = GWZ_CHI2_Q99_TABLE;
// host-side accessors of the shared algorithm (GWZ_HD only because the templates they plug into are __host__ __device__)
struct HostLevels {
    double* p;
    GWZ_HD double& operator()(int level, int field) const { return p[level * gwz::BLK_FIELDS + field]; }
};
struct HostSeries {
    const double* times;
    const double* values;
    GWZ_HD void operator()(uint64_t k, double& tau, double& val) const {
        tau = times[k];
        val = values[k];
    }
};
struct HostSink {
    double* out;
    GWZ_HD void operator()(double v) { *out++ = v; }
};
}  // namespace

extern "C" int gwz_resample(double* result, size_t len, const double* times, const double* values, size_t n) {
    if (!result || !times || !values) return GWZ_ERR_INVALID;
    // state.jl:195-197 throws ArgumentError on a length mismatch; with one n here the analogue is n == 0
    if (n == 0 || len == 0) return GWZ_ERR_INVALID;
    double total_time = 0.0;
    for (size_t j = 0; j < n; ++j) total_time += times[j];  // state.jl:203
    gwz::resample_stream((uint64_t)n, (uint64_t)len, total_time, HostSeries{times, values}, HostSink{result});
    return GWZ_OK;
}

extern "C" int gwz_blocking_analysis(const double* v, size_t n, gwz_blocking_result* out) {
    if (!v || !out) return GWZ_ERR_INVALID;
    const double nan = std::numeric_limits<double>::quiet_NaN();
    out->mean = 0.0;
    out->err = out->err_err = out->p_cov = nan;
    out->k = -1;
    out->blocks = 0;
    if (n < 1) return GWZ_OK;  // "Not enough data for blocking analysis"
    if (n == 1) {
        out->mean = v[0];
        out->k = 0;
        out->blocks = 1;
        return GWZ_OK;
    }
    double centre = 0.0;
    for (size_t i = 0; i < n; ++i) centre += v[i];
    centre /= (double)n;
    const int levels = gwz::blocking_levels(n);
    std::vector<double> store((size_t)levels * gwz::BLK_FIELDS, 0.0);
    gwz::BlockingStream<HostLevels> bs(HostLevels{store.data()}, levels, centre);
    for (size_t i = 0; i < n; ++i) bs.push(v[i]);
    gwz::BlockingOut o;
    bs.finalize((uint64_t)n, kChi2Q99, o);
    out->mean = o.mean;
    out->err = o.err;
    out->err_err = o.err_err;
    out->p_cov = o.p_cov;
    out->k = o.k;
    out->blocks = o.blocks;
    return GWZ_OK;
}
