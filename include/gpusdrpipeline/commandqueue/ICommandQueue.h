/* Forwarding header: the reference's include path for declarations that live in <gpusdrpipeline/abi/queues.h>. */
#ifndef GPUSDRPIPELINE_FWD_COMMANDQUEUE_ICOMMANDQUEUE_H
#define GPUSDRPIPELINE_FWD_COMMANDQUEUE_ICOMMANDQUEUE_H
#include <gpusdrpipeline/abi/queues.h>
#endif
