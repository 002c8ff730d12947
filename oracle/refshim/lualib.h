/* refshim stand-in for lualib.h (TEST INFRASTRUCTURE; see refshim.h) */
#include "refshim.h"
