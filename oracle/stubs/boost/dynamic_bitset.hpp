// Stand-in for <boost/dynamic_bitset.hpp> (boost headers are not installed in this image).
// The PairHMM path only ever *names* the type in a typedef (reference genotype/forward.h:71,
// read_record.h clear_informative_set); it never creates or touches one.  Used solely when
// oracle/Makefile compiles the reference's translation units in place.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <memory>
namespace boost
{
template <typename Block = unsigned long, typename Allocator = std::allocator<Block>>
class dynamic_bitset
{
public:
    dynamic_bitset() = default;
    ~dynamic_bitset() = default;
};
}  // namespace boost
