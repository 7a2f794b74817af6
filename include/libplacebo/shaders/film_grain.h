/* Forwarding header, see ../../libplacebo_b200.h (<libplacebo/shaders/film_grain.h>). */
#include "../../libplacebo_b200.h"
