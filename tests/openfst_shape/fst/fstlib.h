#include "fst/fst.h"
