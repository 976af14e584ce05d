#include "host_util.h"

extern "C" int b200_version(void) { return 2; }
extern "C" int b200_sm_count(void) { return b200::device_sm_count(); }
