#include "clshim.h"
#include "kernels/nbody/buildtree.cl"
