#include "clshim.h"
#include "kernels/nbody/summarizetree.cl"
