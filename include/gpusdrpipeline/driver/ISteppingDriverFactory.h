/* Forwarding header: the reference's include path for declarations that live in <gpusdrpipeline/abi/drivers.h>. */
#ifndef GPUSDRPIPELINE_FWD_DRIVER_ISTEPPINGDRIVERFACTORY_H
#define GPUSDRPIPELINE_FWD_DRIVER_ISTEPPINGDRIVERFACTORY_H
#include <gpusdrpipeline/abi/drivers.h>
#endif
