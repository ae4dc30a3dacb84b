/* forwarding header: see helper/hemoCellStretch.h */
#include "hemocell.h"
