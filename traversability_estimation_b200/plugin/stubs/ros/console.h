// Stand-in for <ros/console.h>, used only when ROS is not installed (this build image).
#pragma once
#include <cstdio>
#define ROS_ERROR(...) do { std::fprintf(stderr, "[ERROR] " __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define ROS_WARN(...) do { std::fprintf(stderr, "[WARN] " __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define ROS_INFO(...) do { } while (0)
#define ROS_DEBUG(...) do { } while (0)
