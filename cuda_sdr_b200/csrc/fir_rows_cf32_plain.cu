// Instantiations of rowsKernel for ELEM=kElemComplex, MIX=false (see fir.cu: rowsKernelFor).
#include "fir_dispatch.h"
#include "fir_kernels.cuh"

namespace b200sdr {
#define ROWS_PAIR(MP) rowsKernel<kElemComplex, false, MP, 1>, rowsKernel<kElemComplex, false, MP, (MP <= 4 ? 4 : 2)>
const FirKernel kRowsCf32Plain[16] = {ROWS_PAIR(1), ROWS_PAIR(2), ROWS_PAIR(3), ROWS_PAIR(4), ROWS_PAIR(5), ROWS_PAIR(6), ROWS_PAIR(7), ROWS_PAIR(8)};
// M = 9..32 with long rows (D >= 16): 16 or 32 partial sums per row stay in registers -- [MP = 16: 1 row, 2 rows][MP = 32: 1 row]
const FirKernel kRowsCf32PlainWide[3] = {rowsKernel<kElemComplex, false, 16, 1>, rowsKernel<kElemComplex, false, 16, 2>, rowsKernel<kElemComplex, false, 32, 1>};
}  // namespace b200sdr
