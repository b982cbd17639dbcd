// hwy_abi.h — error reporting / launch counting shared by the library's translation units.
#pragma once
namespace hwy_abi {
int fail(const char* fmt, const char* detail);   // sets hwy_last_error(), returns 1
int check_launch(const char* what);              // counts the launch, returns 1 on a CUDA error
}  // namespace hwy_abi
