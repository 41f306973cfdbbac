/* TEST-ONLY stand-in for Wolfram's WolframLibrary.h (not present in this image; the real header ships with the
 * Wolfram Engine).  It declares just the subset of the LibraryLink C interface that
 * bayesianinference_b200/wl/librarylink_shim.c uses, with the shapes documented in the LibraryLink user guide, so the
 * shim can be syntax- and link-checked here.  Never used to build the product. */
#ifndef WOLFRAMLIBRARY_STUB_H
#define WOLFRAMLIBRARY_STUB_H
#include <stdint.h>
typedef int64_t mint;
typedef double mreal;
typedef struct st_MTensor *MTensor;
typedef union { mint *integer; mreal *real; MTensor *tensor; char **utf8string; } MArgument;
#define MArgument_getInteger(a) (*((a).integer))
#define MArgument_getReal(a) (*((a).real))
#define MArgument_getMTensor(a) (*((a).tensor))
#define MArgument_setInteger(a, v) ((*((a).integer)) = (v))
#define MArgument_setMTensor(a, v) ((*((a).tensor)) = (v))
#define MArgument_setUTF8String(a, v) ((*((a).utf8string)) = (v))
typedef struct st_WolframLibraryData {
    int (*MTensor_new)(mint, mint, mint const *, MTensor *);
    void (*MTensor_free)(MTensor);
    mint (*MTensor_getRank)(MTensor);
    mint const *(*MTensor_getDimensions)(MTensor);
    mint (*MTensor_getFlattenedLength)(MTensor);
    mint *(*MTensor_getIntegerData)(MTensor);
    mreal *(*MTensor_getRealData)(MTensor);
} *WolframLibraryData;
#define WolframLibraryVersion 7
#define MType_Integer 2
#define MType_Real 3
enum { LIBRARY_NO_ERROR = 0, LIBRARY_TYPE_ERROR, LIBRARY_RANK_ERROR, LIBRARY_DIMENSION_ERROR, LIBRARY_NUMERICAL_ERROR,
       LIBRARY_MEMORY_ERROR, LIBRARY_FUNCTION_ERROR, LIBRARY_VERSION_ERROR };
#define DLLEXPORT __attribute__((visibility("default")))
#endif
