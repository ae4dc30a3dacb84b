/* mechanics/cellMechanics.h of the reference: the CellMechanics base class lives in hemocell.h here */
#pragma once
#include "hemocell.h"
