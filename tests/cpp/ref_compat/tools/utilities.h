// tests/cpp/ref_compat/tools/utilities.h -- the reference's unit tests include
// "tools/utilities.h" (make_unique, saturate) without using it; nothing to provide.
#pragma once
#include <memory>
