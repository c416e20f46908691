/* What CMake generates from the reference's src/config.h.in with WITH_CUDA=ON WITH_TESTS=ON
 * (CMakeLists.txt:3-4, src/CMakeLists.txt:2-12).  TEST INFRASTRUCTURE (oracle/_ref build only). */
#ifndef CONFIG_HPP
#define CONFIG_HPP
#define TESTING_ENABLED
#define CUDA_ENABLED
#endif
