/* Forwarding header, see ../../libplacebo_b200.h (<libplacebo/shaders/sampling.h>). */
#include "../../libplacebo_b200.h"
