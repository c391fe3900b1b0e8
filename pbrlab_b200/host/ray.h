// forwards to api-types.h (struct Ray)
#pragma once
#include "api-types.h"
