/* Forwarding header, see ../../libplacebo_b200.h (<libplacebo/shaders/colorspace.h>). */
#include "../../libplacebo_b200.h"
