// Core include of the (absent) Veles platform, written for the reference-arm shim: the
// reference's cuda/*.cu expect ``dtype``, ``SIGN``, ``MIN``/``MAX`` and a double atomicAdd.
#ifndef _VELES_DEFINES_CU_
#define _VELES_DEFINES_CU_

#ifndef FLT_MAX
#define FLT_MAX 3.402823466e+38f
#endif
#ifndef DBL_MAX
#define DBL_MAX 1.7976931348623158e+308
#endif
#ifndef INT_MAX
#define INT_MAX 2147483647
#endif

#define SIGN(x) ((x) ? ((x) > 0 ? 1 : -1) : 0)
#ifndef MIN
#define MIN(a, b) ((a) < (b) ? (a) : (b))
#endif
#ifndef MAX
#define MAX(a, b) ((a) > (b) ? (a) : (b))
#endif

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef unsigned long long ulong;

#endif  // _VELES_DEFINES_CU_
