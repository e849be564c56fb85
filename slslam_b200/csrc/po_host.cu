// placeholder until the PO kernels land (K5/K6)
#include "common_host.h"
extern "C" {
int slslam_po_solve(const slslam_po_desc*, double*, slslam_summary*) { return SLSLAM_ERR_UNSUPPORTED; }
int slslam_po_solve_trace(const slslam_po_desc*, double*, slslam_summary*, double*) { return SLSLAM_ERR_UNSUPPORTED; }
int slslam_po_evaluate(const slslam_po_desc*, const double*, double*, double*, double*, double*) { return SLSLAM_ERR_UNSUPPORTED; }
}
