// TEST INFRASTRUCTURE ONLY.  Prologue of oracle/_ref/ref_host_gen.cpp: the reference's HOST-side arithmetic
// (C++, src/loader/material.cpp, src/core/camera.cpp, src/core/texture.cpp) extracted function by function at
// build time and compiled against the header-only nvmath vendored in the reference tree.  The type aliases are
// the C++ branch of src/shared/binding.h:4-13; VkExtent2D stands in for <vulkan/vulkan_core.h>.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <nvmath/nvmath.h>
using std::string;
using std::vector;
using vec2 = nvmath::vec2f;
using vec3 = nvmath::vec3f;
using vec4 = nvmath::vec4f;
using mat4 = nvmath::mat4f;
using uint = unsigned int;
struct VkExtent2D {
  uint32_t width, height;
};
