// forwards to ../api-types.h (struct Attribute, struct CurveAttribute)
#pragma once
#include "../api-types.h"
