/* refshim stand-in for lauxlib.h (TEST INFRASTRUCTURE; see refshim.h) */
#include "refshim.h"
