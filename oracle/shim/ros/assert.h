// oracle/shim: the ROS assertion / logging macros the factor sources use (TEST INFRASTRUCTURE)
#pragma once
#include <cassert>
#include <cstdio>
#include <cstdlib>
#define ROS_ASSERT(x) assert(x)
#define ROS_ASSERT_MSG(x, ...) assert(x)
#define ROS_BREAK() std::abort()
#define ROS_WARN(...) std::fprintf(stderr, __VA_ARGS__)
#define ROS_INFO(...) ((void)0)
#define ROS_DEBUG(...) ((void)0)
#define ROS_ERROR(...) std::fprintf(stderr, __VA_ARGS__)
#define ROS_WARN_STREAM(x) ((void)0)
#define ROS_INFO_STREAM(x) ((void)0)
