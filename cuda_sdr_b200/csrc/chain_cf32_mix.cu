// Instantiations of chainKernel for ELEM=kElemComplex, MIX=true (see chain_launch.cu: chainKernelFor).
#include "chain_dispatch.h"
#include "chain_kernels.cuh"

namespace b200sdr {
#define CHAIN_SET(MP)                                                                                       \
  chainKernel<kElemComplex, true, MP, 2, 0>, chainKernel<kElemComplex, true, MP, 2, 0>,                                    \
      chainKernel<kElemComplex, true, MP, (MP <= 4 ? 4 : 2), 0>, chainKernel<kElemComplex, true, MP, (MP <= 4 ? 4 : 2), 0>
const ChainKernel kChainCf32Mix[32] = {CHAIN_SET(1), CHAIN_SET(2), CHAIN_SET(3), CHAIN_SET(4),
                                CHAIN_SET(5), CHAIN_SET(6), CHAIN_SET(7), CHAIN_SET(8)};
}  // namespace b200sdr
