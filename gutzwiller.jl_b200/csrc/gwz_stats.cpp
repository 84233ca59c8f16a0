// gwz_stats.cpp -- error bars for recorded series: resample! (src/kinetic-vqmc/state.jl:194-231)
// and Rimu.StatsTools.blocking_analysis (Flyvbjerg-Petersen reblocking with the automated
// M-test of Jonsson, PRE 98, 043304 (2018); used at state.jl:53,273).  Host code, like in the
// reference: these run once per walker on series that were recorded by the GPU kernel.
#include "gutzwiller_b200.h"

#include <cmath>
#include <limits>
#include <vector>

extern "C" int gwz_resample(double* result, size_t len, const double* times, const double* values, size_t n) {
    if (!result || !times || !values) return GWZ_ERR_INVALID;
    // state.jl:195-197 throws ArgumentError on a length mismatch; with one n here the analogue is n == 0
    if (n == 0 || len == 0) return GWZ_ERR_INVALID;
    double total_time = 0.0;
    for (size_t j = 0; j < n; ++j) total_time += times[j];
    // Each input interval k occupies len*times[k]/total_time output slots; slot i covers (i-1, i].
    double t = 0.0;    // end of the intervals consumed so far, in slot units
    double mu = 0.0;   // integral of the series over the part of slot i seen so far
    double val = 0.0;  // value of the interval that is currently spilling over
    size_t k = 1, i = 1;
    for (;;) {
        while (t >= (double)i && i <= len) {  // whole slots covered by the spilling interval
            result[i - 1] = val;
            ++i;
            mu -= val;
        }
        while (t < (double)i && k <= n) {  // consume intervals until slot i is full
            val = values[k - 1];
            const double step = (double)len * times[k - 1] / total_time;
            const double room = (double)i - t;
            mu += val * (step < room ? step : room);
            t += step;
            ++k;
        }
        if (i > len) break;
        result[i - 1] = mu;
        mu = (t - (double)i) * val;  // carry the overshoot into the next slot
        ++i;
    }
    return GWZ_OK;
}

namespace {
// chi-square 0.99 quantiles for 1..30 degrees of freedom (alpha = 0.01)
const double kChi2Q99[30] = {
    6.634896601021213,  9.210340371976180,  11.344866730144370, 13.276704135987622, 15.086272469388987,
    16.811893829770927, 18.475306906582357, 20.090235029663233, 21.665994333461924, 23.209251158954356,
    24.724970311318277, 26.216967305535853, 27.688249610457049, 29.141237740672796, 30.577914166892494,
    31.999926908815176, 33.408663605004612, 34.805305734705072, 36.190869129270048, 37.566234786625053,
    38.932172683516072, 40.289360437593864, 41.638398118858476, 42.979820139351631, 44.314104896219156,
    45.641682666283153, 46.962942124751443, 48.278235770315498, 49.587884472898835, 50.892181311517092};
}  // namespace

extern "C" int gwz_blocking_analysis(const double* v_in, size_t n, gwz_blocking_result* out) {
    if (!v_in || !out) return GWZ_ERR_INVALID;
    const double nan = std::numeric_limits<double>::quiet_NaN();
    out->mean = 0.0;
    out->err = out->err_err = out->p_cov = nan;
    out->k = -1;
    out->blocks = 0;
    if (n < 1) return GWZ_OK;
    if (n == 1) {
        out->mean = v_in[0];
        out->k = 0;
        out->blocks = 1;
        return GWZ_OK;
    }
    int levels = 0;
    while (((size_t)1 << (levels + 1)) <= n) ++levels;
    std::vector<double> v(v_in, v_in + n);
    std::vector<double> se(levels), see(levels), mj(levels);
    std::vector<int64_t> blocks(levels);
    size_t len = n;
    double mean0 = 0.0;
    for (int s = 0; s < levels; ++s) {
        long double sum = 0;
        for (size_t i = 0; i < len; ++i) sum += v[i];
        const double mean = (double)(sum / (long double)len);
        if (s == 0) mean0 = mean;
        long double ss = 0, gam = 0;
        for (size_t i = 0; i < len; ++i) {
            const long double d = (long double)v[i] - mean;
            ss += d * d;
            if (i + 1 < len) gam += d * ((long double)v[i + 1] - mean);
        }
        const double var_unc = (double)(ss / (long double)len);
        const double var_cor = len > 1 ? (double)(ss / (long double)(len - 1)) : nan;
        const double gamma = (double)(gam / (long double)len);
        blocks[s] = (int64_t)len;
        se[s] = std::sqrt(var_cor / (double)len);
        see[s] = se[s] / std::sqrt(2.0 * (double)(len - 1));
        const double q = (double)(len - 1) * var_unc / ((double)len * (double)len) + gamma;
        mj[s] = var_unc > 0.0 ? (double)len * q * q / (var_unc * var_unc) : 0.0;
        const size_t half = len / 2;
        for (size_t i = 0; i < half; ++i) v[i] = 0.5 * (v[2 * i] + v[2 * i + 1]);
        len = half;
    }
    // M_k = sum_{j>=k} mj[j]; accept the first level whose M_k is below the chi-square quantile
    std::vector<double> M(levels);
    double tail = 0.0;
    for (int s = levels - 1; s >= 0; --s) {
        tail += mj[s];
        M[s] = tail;
    }
    int kk = -1;
    for (int s = 0; s < levels; ++s) {
        const int dof = levels - s;
        const double q = dof <= 30 ? kChi2Q99[dof - 1] : kChi2Q99[29] + 1.3 * (dof - 30);
        if (M[s] < q) {
            kk = s;
            break;
        }
    }
    out->mean = mean0;
    if (kk >= 0) {
        out->err = se[kk];
        out->err_err = see[kk];
        out->p_cov = 0.0;
        out->k = kk + 1;
        out->blocks = blocks[kk];
    }
    return GWZ_OK;
}
