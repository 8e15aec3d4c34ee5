// malloc-backed stand-in for boost::pool<> (see poolfwd.hpp in this directory). TEST INFRASTRUCTURE.
#ifndef IPT_B200_ORACLE_POOL_SHIM
#define IPT_B200_ORACLE_POOL_SHIM
#include "poolfwd.hpp"
#include <cstdlib>
namespace boost {
struct default_user_allocator_new_delete {};
template <typename UserAllocator> class pool {
    std::size_t chunk_;
public:
    explicit pool(std::size_t requested_size) : chunk_(requested_size) {}
    void* malloc() { return std::malloc(chunk_); }
    void free(void* p) { std::free(p); }
};
}
#endif
