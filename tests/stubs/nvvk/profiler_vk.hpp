// TEST STUB, see resourceallocator_vk.hpp
#pragma once
namespace nvvk {
class ProfilerVK
{
};
}  // namespace nvvk
