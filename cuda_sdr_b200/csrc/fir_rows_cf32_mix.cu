// Instantiations of rowsKernel for ELEM=kElemComplex, MIX=true (see fir.cu: rowsKernelFor).
#include "fir_dispatch.h"
#include "fir_kernels.cuh"

namespace b200sdr {
#define ROWS_PAIR(MP) rowsKernel<kElemComplex, true, MP, 1>, rowsKernel<kElemComplex, true, MP, (MP <= 4 ? 4 : 2)>
const FirKernel kRowsCf32Mix[16] = {ROWS_PAIR(1), ROWS_PAIR(2), ROWS_PAIR(3), ROWS_PAIR(4), ROWS_PAIR(5), ROWS_PAIR(6), ROWS_PAIR(7), ROWS_PAIR(8)};
}  // namespace b200sdr
