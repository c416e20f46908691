/* stand-in, see sam.h */
#include "sam.h"
