/* Forwarding header: the libplacebo API surface of the B200 hot path lives in one file
 * (include/libplacebo_b200.h); this lets sources written against <libplacebo/renderer.h> compile
 * unchanged. */
#include "../libplacebo_b200.h"
