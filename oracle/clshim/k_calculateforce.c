#include "clshim.h"
#include CALCULATEFORCE_CL
