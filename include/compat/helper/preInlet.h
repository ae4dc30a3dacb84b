#include "hemocell.h"
