// tests/cpp/ref_compat/fbstab/test/ocp_generator.h -- the reference keeps its problem
// generator under fbstab/test/; the facade ships the same class as fbstab/ocp_generator.h.
#pragma once
#include "fbstab/ocp_generator.h"
