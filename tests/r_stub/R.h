/* tests/r_stub/R.h -- TEST INFRASTRUCTURE, see Rinternals.h in this directory */
#ifndef RSTUB_R_H
#define RSTUB_R_H
#include <stdio.h>
#include <stdlib.h>
#endif
