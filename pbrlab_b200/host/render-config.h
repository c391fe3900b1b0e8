// forwards to api-types.h (struct RenderConfig)
#pragma once
#include "api-types.h"
