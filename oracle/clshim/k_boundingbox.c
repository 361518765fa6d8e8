#include "clshim.h"
#include "kernels/nbody/boundingbox.cl"
