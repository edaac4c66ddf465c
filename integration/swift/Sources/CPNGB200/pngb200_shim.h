/* One include for the module map: the real header lives in <pngb200 repo>/include (keep a copy or a symlink
 * beside this file, or pass -Xcc -I<pngb200 repo>/include). */
#include "pngb200.h"
