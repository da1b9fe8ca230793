/* Forced-include prelude for compiling the UNMODIFIED reference headers with
 * g++ 13 (test infrastructure only, oracle/).
 *
 * 1. GCC 13's libstdc++ no longer includes <functional>/<cstddef> transitively;
 *    apf/apf/convolver.h:792 uses std::bind and apf/apf/fftwtools.h:85 uses
 *    ptrdiff_t without including them.
 * 2. apf/apf/stringtools.h:66-76 (`convert`) does `input >> result >> std::ws`
 *    and then requires `!fail() && eof()`.  With a current libstdc++ `>> std::ws`
 *    on a stream that is already at EOF sets failbit, so every
 *    apf::parameter_map::get<T>() throws.  Non-template overloads declared here,
 *    BEFORE stringtools.h is parsed, are preferred by overload resolution inside
 *    S2A() and restore the intended behaviour (value followed only by optional
 *    whitespace).  No reference source is modified or copied.
 */
#ifndef SSR_B200_ORACLE_PRELUDE_H
#define SSR_B200_ORACLE_PRELUDE_H
#ifdef __cplusplus
#include <cstddef>
#include <functional>
#include <istream>
#include <string>

namespace apf { namespace str {

#define SSR_B200_CONVERT_OVERLOAD(T) \
inline bool convert(std::istream& input, T& output) \
{ \
  T result{}; \
  input >> result; \
  if (input.fail()) return false; \
  if (!input.eof()) input >> std::ws; \
  if (!input.eof()) return false; \
  output = result; \
  return true; \
}

SSR_B200_CONVERT_OVERLOAD(int)
SSR_B200_CONVERT_OVERLOAD(unsigned int)
SSR_B200_CONVERT_OVERLOAD(long)
SSR_B200_CONVERT_OVERLOAD(unsigned long)
SSR_B200_CONVERT_OVERLOAD(float)
SSR_B200_CONVERT_OVERLOAD(double)
SSR_B200_CONVERT_OVERLOAD(std::string)

#undef SSR_B200_CONVERT_OVERLOAD

}}  // namespace apf::str
#endif
#endif
