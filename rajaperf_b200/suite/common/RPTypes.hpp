// RPTypes.hpp -- scalar types of the suite (reference: common/RPTypes.hpp:49-107).
#pragma once
#include <cstddef>
#include <cstdint>

namespace rajaperf {
using Real_type = double;            // RPTypes.hpp:107
using Real_ptr = Real_type*;
using Index_type = std::ptrdiff_t;   // RAJA::Index_type (tpl/RAJA/include/RAJA/util/types.hpp:179)
using Int_type = int;                // RPTypes.hpp:81 -- halo index lists
using Int_ptr = Int_type*;
using Checksum_type = long double;   // RPTypes.hpp: Checksum_type
using RepIndex_type = volatile int;  // RPTypes.hpp:49
using Size_type = std::size_t;
}  // namespace rajaperf
