/* refshim stand-in for lua.h (TEST INFRASTRUCTURE; see refshim.h) */
#include "refshim.h"
