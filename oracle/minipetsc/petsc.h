#include "minipetsc.h"
