// Stand-in for the reference's CMake-generated generated/defines.h (CMakeLists.txt:204-229).
// Written for the oracle/_ref build recipe; selects the CUDA (float3) struct layout.
#ifndef SOLR_B200_REF_DEFINES_H
#define SOLR_B200_REF_DEFINES_H
#include <cstddef>
#ifndef USE_CUDA
#define USE_CUDA 1
#endif
static const char* DEFAULT_KERNEL_FILENAME = "";
static const char* DEFAULT_MEDIA_FOLDER = ".";
#endif
