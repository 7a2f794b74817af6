/* Forwarding header, see ../../libplacebo_b200.h (<libplacebo/shaders/dithering.h>). */
#include "../../libplacebo_b200.h"
