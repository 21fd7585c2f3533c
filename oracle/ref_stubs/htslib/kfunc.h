/* oracle/ref_stubs: declaration of htslib's kf_erfc; the definition is the oracle's restatement
 * (oracle_stats.c:oracle_kf_erfc) via ref_glue.c. TEST INFRASTRUCTURE ONLY. */
#ifndef STUB_HTSLIB_KFUNC_H
#define STUB_HTSLIB_KFUNC_H
double kf_erfc(double x);
#endif
