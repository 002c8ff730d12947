/* refshim stand-in for THC.h (TEST INFRASTRUCTURE; see refshim.h) */
#include "refshim.h"
