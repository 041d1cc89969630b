/* TEST INFRASTRUCTURE ONLY (oracle/): the four libpng typedefs pngfuncs.h mentions
   (pngfuncs.cpp itself is never compiled: nothing references it). */
#ifndef DFTB200_ORACLE_SHIM_PNG
#define DFTB200_ORACLE_SHIM_PNG
#include <stddef.h>
typedef struct png_struct_def* png_structp;
typedef unsigned char* png_bytep;
typedef size_t png_size_t;
typedef struct png_info_def* png_infop;
#endif
