// hm_cpu_guard.cpp — compiled WITHOUT the -march flags of the rest of the library (frontend/Makefile): the decoder objects are
// built for x86-64-v3 (AVX2, BMI1/2, LZCNT, MOVBE, FMA: every host a B200 sits in), and this is the one function that may run on
// anything and say so, instead of the process dying of an illegal instruction somewhere inside the parser.
extern "C" int hm_cpu_is_x86_64_v3(void)
{
#if defined(__x86_64__)
  __builtin_cpu_init();
  return __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi") && __builtin_cpu_supports("bmi2") &&
         __builtin_cpu_supports("fma") && __builtin_cpu_supports("popcnt");
#else
  return 1;
#endif
}
