/* forwarding header: the case files of the reference include "cellInfo.h"; its declarations live in hemocell.h here */
#include "hemocell.h"
