// Fused red+black Gauss-Seidel pass (placeholder until the streaming kernel lands).
#include "common.cuh"
extern "C" {
int b200mg_gsrb2_abec (int, const b200mg_tile*, const b200mg_box*, const b200mg_fab*, const b200mg_fab*, const b200mg_fab*, const b200mg_fab*,
                       const b200mg_fab*, const b200mg_fab*, const b200mg_fab*, const b200mg_fab*, const b200mg_ifab*,
                       double, double, double, double, int, int, cudaStream_t) { return int(cudaErrorNotSupported); }
int b200mg_gsrb2_poisson (int, const b200mg_tile*, const b200mg_box*, const b200mg_fab*, const b200mg_fab*, const b200mg_fab*,
                          const b200mg_fab*, const b200mg_ifab*, double, double, double, int, int, cudaStream_t) { return int(cudaErrorNotSupported); }
}
