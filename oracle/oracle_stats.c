/*
 * oracle/oracle_stats.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Site statistics restated from the reference:
 *   calc_vdb        src/ext/bcftools/bcftools-ext.c:51-111
 *   calc_mwu_biasZ  src/ext/bcftools/bcftools-ext.c:144-212 (raer calls it with left_only=0, do_Z=1,
 *                   src/plp.c:338; the exact-table branch is therefore unreachable and not restated)
 *   calc_sor        src/plp_utils.c:375-387
 *   kf_erfc         htslib kfunc.c (not vendored): Applied Statistics algorithm AS66, 2nd form.
 * tests/test_oracle_stats.py cross-checks these against the reference's own bcftools-ext.c built
 * into oracle/_ref/ (see oracle/Makefile) when /root/reference is present.
 */
#include "oracle.h"

#include <math.h>

double oracle_kf_erfc(double x) {
    static const double p[7] = {220.2068679123761, 221.2135961699311, 112.0792914978709, 33.912866078383,
                                6.37396220353165,  .7003830644436881, .03526249659989109};
    static const double q[8] = {440.4137358247522, 793.8265125199484, 637.3336333788311, 296.5642487796737,
                                86.78073220294608, 16.06417757920695, 1.755667163182642, .08838834764831844};
    double z = fabs(x) * M_SQRT2;
    if (z > 37.) return x > 0. ? 0. : 2.;
    double e = exp(z * z * -.5), r;
    if (z < 10. / M_SQRT2) {
        double num = ((((((p[6] * z + p[5]) * z + p[4]) * z + p[3]) * z + p[2]) * z + p[1]) * z + p[0]);
        double den = (((((((q[7] * z + q[6]) * z + q[5]) * z + q[4]) * z + q[3]) * z + q[2]) * z + q[1]) * z + q[0]);
        r = e * num / den;
    } else {
        r = e / 2.506628274631001 / (z + 1. / (z + 2. / (z + 3. / (z + 4. / (z + .65)))));
    }
    return x > 0. ? 2. * r : 2. * (1. - r);
}

double oracle_calc_vdb(const int* pos, int npos) {
    const int readlen = 100;
    static const float prm[15][3] = {{3, 0.079f, 18},    {4, 0.09f, 19.8f},  {5, 0.1f, 20.5f},   {6, 0.11f, 21.5f},
                                     {7, 0.125f, 21.6f}, {8, 0.135f, 22},    {9, 0.14f, 22.2f},  {10, 0.153f, 22.3f},
                                     {15, 0.19f, 22.8f}, {20, 0.22f, 23.2f}, {30, 0.26f, 23.4f}, {40, 0.29f, 23.5f},
                                     {50, 0.35f, 23.65f}, {100, 0.5f, 23.7f}, {200, 0.7f, 23.7f}};
    int i, dp = 0;
    float mean_pos = 0, mean_diff = 0;
    for (i = 0; i < npos; ++i) {
        if (!pos[i]) continue;
        dp += pos[i];
        mean_pos += pos[i] * i; /* int product converted to float, float accumulate */
    }
    if (dp < 2) return HUGE_VAL;
    mean_pos /= dp;
    for (i = 0; i < npos; ++i) {
        if (!pos[i]) continue;
        /* upstream: float += int * fabs(int - float): the product is evaluated in double */
        mean_diff += pos[i] * fabs(i - mean_pos);
    }
    mean_diff /= dp;
    int ipos = mean_diff;
    if (dp == 2) return (2 * readlen - 2 * (ipos + 1) - 1) * (ipos + 1) / (readlen - 1) / (readlen * 0.5);
    if (dp >= 200) {
        i = 15;
    } else {
        for (i = 0; i < 15; ++i)
            if (prm[i][0] >= dp) break;
    }
    float pshift, pscale;
    if (i == 15) {
        pscale = prm[14][1];
        pshift = prm[14][2];
    } else if (i > 0 && prm[i][0] != dp) {
        pscale = (prm[i - 1][1] + prm[i][1]) * 0.5;
        pshift = (prm[i - 1][2] + prm[i][2]) * 0.5;
    } else {
        pscale = prm[i][1];
        pshift = prm[i][2];
    }
    return 0.5 * oracle_kf_erfc(-(mean_diff - pshift) * pscale);
}

double oracle_calc_mwu_biasZ(const int* a, const int* b, int n, int left_only, int do_Z) {
    int i;
    int64_t t = 0;
    for (i = 0; i < n; ++i)
        if (b[i]) break;
    int b_empty = (i == n);
    int e = 0, l = 0, na = 0, nb = 0;
    if (b_empty) {
        for (i = n - 1; i >= 0; --i) {
            na += a[i];
            t += (a[i] * a[i] - 1) * a[i];
        }
    } else {
        for (i = n - 1; i >= 0; --i) {
            e += a[i] * b[i];
            l += a[i] * nb;
            na += a[i];
            nb += b[i];
            int p = a[i] + b[i];
            t += (p * p - 1) * p;
        }
    }
    if (!na || !nb) return HUGE_VAL;
    double U = l + e * 0.5;
    double m = na * nb / 2.0;
    double var2 = (na * nb) / 12.0 * ((na + nb + 1) - t / (double)((na + nb) * (na + nb - 1)));
    if (var2 <= 0) return do_Z ? 0 : 1;
    if (do_Z) return (U - m) / sqrt(var2);
    if (left_only && U > m) return HUGE_VAL;
    /* raer never reaches here (do_Z == 1): normal approximation only */
    return exp(-0.5 * (U - m) * (U - m) / var2);
}

double oracle_calc_sor(int fwd_ref, int rev_ref, int fwd_alt, int rev_alt) {
    double t00 = (double)fwd_ref + 1, t01 = (double)rev_ref + 1;
    double t11 = (double)fwd_alt + 1, t10 = (double)rev_alt + 1;
    double ratio = (t00 / t01) * (t11 / t10) + (t01 / t00) * (t10 / t11);
    double ref_ratio = fmin(t00, t01) / fmax(t00, t01);
    double alt_ratio = fmin(t10, t11) / fmax(t10, t11);
    return log(ratio) + log(ref_ratio) - log(alt_ratio);
}
