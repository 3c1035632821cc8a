/* Hand-written (see config.h in this directory). */
#define _FILE_OFFSET_BITS 64
#define GMX_FAHCORE 0
#define TMPI_WAIT_FOR_NO_ONE 0
