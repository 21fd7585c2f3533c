// raer_sc.cu -- single-cell pileup path (placeholder until the device path lands below).
#include "raer_b200.h"
#include "raer_host.h"
extern "C" int raer_scpileup(const raer_scpileup_args* a) {
    (void)a;
    raer::set_error("raer_scpileup: device path not built yet");
    return RAER_ERR_LIMIT;
}
