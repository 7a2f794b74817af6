/* Forwarding header, see ../../libplacebo_b200.h (<libplacebo/utils/upload.h>). */
#include "../../libplacebo_b200.h"
