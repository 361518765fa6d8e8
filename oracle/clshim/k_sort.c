#include "clshim.h"
#include "kernels/nbody/sort.cl"
