/* Forwarding header: the reference's include path for declarations that live in <gpusdrpipeline/abi/buffers.h>. */
#ifndef GPUSDRPIPELINE_FWD_BUFFERS_ICUDABUFFERCOPIERFACTORY_H
#define GPUSDRPIPELINE_FWD_BUFFERS_ICUDABUFFERCOPIERFACTORY_H
#include <gpusdrpipeline/abi/buffers.h>
#endif
