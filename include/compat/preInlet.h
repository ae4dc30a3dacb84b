/* helper/preInlet.h of the reference: everything lives in hemocell.h here */
#include "hemocell.h"
