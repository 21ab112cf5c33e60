/* mock: see Rinternals.h in this directory */
#include <stdlib.h>
