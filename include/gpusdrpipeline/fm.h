/* Forwarding header: the reference's include path for declarations that live in <gpusdrpipeline/abi/core.h>. */
#ifndef GPUSDRPIPELINE_FWD_FM_H
#define GPUSDRPIPELINE_FWD_FM_H
#include <gpusdrpipeline/abi/core.h>
#endif
