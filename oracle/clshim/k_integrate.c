#include "clshim.h"
#include "kernels/nbody/integrate.cl"
