// TEST INFRASTRUCTURE ONLY: the handful of Win32 names the reference's headers
// (framework.h, platform_common.h) mention, so they parse on Linux.  No behaviour.
#pragma once
#ifndef RTB_EMU_WINDOWS_H
#define RTB_EMU_WINDOWS_H
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <cmath>
#include <typeinfo>
#include <algorithm>
#include <string>
#include <fstream>
#include <sstream>
#include <unordered_map>
#include <vector>
typedef void* HWND;
enum { VK_UP = 0x26, VK_DOWN = 0x28, VK_LEFT = 0x25, VK_RIGHT = 0x27, VK_RETURN = 0x0D, VK_ESCAPE = 0x1B, VK_F1 = 0x70 };
#ifndef min
#define min(a, b) (((a) < (b)) ? (a) : (b))
#endif
#ifndef max
#define max(a, b) (((a) > (b)) ? (a) : (b))
#endif
#endif
