// forwards to the consolidated stand-in for the un-vendored upstream interfaces (shim_core.h explains)
#include "../shim_core.h"
