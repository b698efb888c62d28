// see stub_core.h (test-only stand-in for the TensorFlow header of the same name)
#pragma once
#include "stub_core.h"
