// Allocator-only stand-in for the boost::pool forward header that dimalit/ipt vendors without
// boost/config.hpp. TEST INFRASTRUCTURE: used only to compile the untouched
// reference sources into oracle/_ref/. boost::pool is used by the reference purely as a fixed-size
// allocator (src/libddf/ddf.cpp:16-56, src/main.cpp:46-49), so no arithmetic depends on it.
#ifndef IPT_B200_ORACLE_POOLFWD_SHIM
#define IPT_B200_ORACLE_POOLFWD_SHIM
#include <cstddef>
namespace boost {
struct default_user_allocator_new_delete;
template <typename UserAllocator = default_user_allocator_new_delete> class pool;
}
#endif
