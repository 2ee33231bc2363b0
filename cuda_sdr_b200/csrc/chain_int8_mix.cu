// Instantiations of chainKernel for ELEM=kElemInt8Complex, MIX=true (see chain_launch.cu: chainKernelFor).
// Per MP: rows per lane {1, 2, 4 (2 when MP > 4)} x route {CUDA cores, int8 tensor cores (int8 input only)}.
#include "chain_dispatch.h"
#include "chain_kernels.cuh"

namespace b200sdr {
#define CHAIN_SET(MP)                                                                       \
  chainKernel<kElemInt8Complex, true, MP, 1, 1>, chainKernel<kElemInt8Complex, true, MP, 1, 2>,               \
      chainKernel<kElemInt8Complex, true, MP, 2, 1>, chainKernel<kElemInt8Complex, true, MP, 2, 2>,           \
      chainKernel<kElemInt8Complex, true, MP, (MP <= 4 ? 4 : 2), 1>, chainKernel<kElemInt8Complex, true, MP, (MP <= 4 ? 4 : 2), 2>
const ChainKernel kChainInt8Mix[48] = {CHAIN_SET(1), CHAIN_SET(2), CHAIN_SET(3), CHAIN_SET(4),
                                CHAIN_SET(5), CHAIN_SET(6), CHAIN_SET(7), CHAIN_SET(8)};
}  // namespace b200sdr
