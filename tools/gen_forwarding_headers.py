#!/usr/bin/env python3
"""Writes the one-line forwarding headers that keep the reference's include paths working
(`#include <gpusdrpipeline/buffers/IBuffer.h>` etc.) on top of the consolidated headers under
include/gpusdrpipeline/abi/.  Re-run after adding a name; the generated files are committed."""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include", "gpusdrpipeline")

GROUPS = {
    "abi/core.h": ["Status.h", "GSDefs.h", "IRef.h", "Result.h", "GSLog.h", "GSErrors.h", "CudaErrors.h", "SampleType.h",
                   "Modulation.h", "am.h", "fm.h", "util/CudaUtil.h", "util/CudaDevicePushPop.h"],
    "abi/buffers.h": ["IMemory.h"] + ["buffers/%s.h" % n for n in (
        "IAllocator IAllocatorFactory IBuffer IBufferCopier IBufferFactory IBufferPool IBufferPoolFactory IBufferRange "
        "IBufferRangeFactory IBufferRangeMutableCapacity IBufferSliceFactory IBufferUtil ICudaAllocatorFactory "
        "ICudaBufferCopierFactory ICudaMemSetFactory IMemSet IRelocatable IRelocatableCudaBufferFactory "
        "IRelocatableResizableBuffer IRelocatableResizableBufferFactory IResizable IResizableBuffer "
        "IResizableBufferFactory").split()],
    "abi/queues.h": ["commandqueue/%s.h" % n for n in
                     "ICommandQueue ICommandQueueFactory ICudaCommandQueue ICudaCommandQueueFactory".split()],
    "abi/nodes.h": ["filters/%s.h" % n for n in
                    "Filter FilterFactories IHackrfSource IPortRemappingSink IPortRemappingSource IReadByteCountMonitor".split()],
    "abi/drivers.h": ["driver/%s.h" % n for n in
                      "IDriver IDriverToDiagram IDriverToDiagramFactory IFilterDriver IFilterDriverFactory ISteppingDriver "
                      "ISteppingDriverFactory".split()],
}

for target, names in GROUPS.items():
    for name in names:
        path = os.path.join(INC, name)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        guard = "GPUSDRPIPELINE_FWD_" + name.replace("/", "_").replace(".", "_").upper()
        with open(path, "w") as f:
            f.write(f"/* Forwarding header: the reference's include path for declarations that live in <gpusdrpipeline/{target}>. */\n"
                    f"#ifndef {guard}\n#define {guard}\n#include <gpusdrpipeline/{target}>\n#endif\n")
print("wrote", sum(len(v) for v in GROUPS.values()), "forwarding headers")
