/* forwarding header: the case files of the reference include "helper/hemocellInit.hh"; its declarations live in hemocell.h here */
#include "hemocell.h"
