/* oracle/ref_glue.c -- TEST INFRASTRUCTURE ONLY: supplies kf_erfc to the reference object code. */
double oracle_kf_erfc(double x);
double kf_erfc(double x) { return oracle_kf_erfc(x); }
