#include "../../include/b200dino.h"
extern "C" const char* b200_version(void) { return "b200dino 0.1.0 sm_100a"; }
