/* The case files of the reference include "palabos3D.h" / "palabos3D.hh" next to "hemocell.h".  Here the
 * Palabos calls they make are served by the plb:: shim inside hemocell.h. */
#ifndef HEMOCELL_PALABOS3D_SHIM_H
#define HEMOCELL_PALABOS3D_SHIM_H
#include "hemocell.h"
#endif
