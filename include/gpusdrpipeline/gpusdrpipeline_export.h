/* Symbol visibility of libgpusdrpipeline.so.  The reference generates this header with CMake's
 * generate_export_header (reference src/CMakeLists.txt:259-273; included by include/gpusdrpipeline/GSDefs.h:20);
 * this repo builds without CMake, so it is written out by hand. */
#ifndef GPUSDRPIPELINE_EXPORT_H
#define GPUSDRPIPELINE_EXPORT_H

#define GS_PUBLIC __attribute__((visibility("default")))
#define GS_PRIVATE __attribute__((visibility("hidden")))

#endif
