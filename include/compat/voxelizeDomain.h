/* forwarding header: the case files of the reference include "helper/voxelizeDomain.h"; its declarations live in hemocell.h here */
#include "hemocell.h"
