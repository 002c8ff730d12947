/* refshim stand-in for luaT.h (TEST INFRASTRUCTURE; see refshim.h) */
#include "refshim.h"
