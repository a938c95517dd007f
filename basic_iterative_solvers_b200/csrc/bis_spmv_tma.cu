// bis_spmv_tma.cu -- SpMV variant 2 (TMA-staged, thread-per-row); placeholder until measured.
#include "bis_device.cuh"

int bis_spmv_tma_try(bis_context *, const bis_matrix *, const double *, int, const void *, int, int) {
    return -1;
}
