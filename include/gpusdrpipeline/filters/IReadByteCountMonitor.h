/* Forwarding header: the reference's include path for declarations that live in <gpusdrpipeline/abi/nodes.h>. */
#ifndef GPUSDRPIPELINE_FWD_FILTERS_IREADBYTECOUNTMONITOR_H
#define GPUSDRPIPELINE_FWD_FILTERS_IREADBYTECOUNTMONITOR_H
#include <gpusdrpipeline/abi/nodes.h>
#endif
