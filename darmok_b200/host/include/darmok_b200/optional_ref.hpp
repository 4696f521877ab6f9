// darmok_b200/optional_ref.hpp -- darmok::OptionalRef<T> for the shim (include/darmok/optional_ref.hpp): what
// SkeletalAnimator::getCurrentState / getCurrentTransition return (include/darmok/skeleton.hpp:283-284).  Inside a darmok build
// its own header is used; stand-alone, the subset below: a nullable reference with pointer-like access.
#pragma once

#if __has_include(<darmok/optional_ref.hpp>)
#include <darmok/optional_ref.hpp>
#else
#include <cstddef>
#include <optional>
#include <stdexcept>
namespace darmok {
template <typename T>
class OptionalRef final {
public:
    OptionalRef() noexcept = default;
    OptionalRef(std::nullopt_t) noexcept {}
    OptionalRef(std::nullptr_t) noexcept {}
    OptionalRef(T* value) noexcept : _value(value) {}
    OptionalRef(T& value) noexcept : _value(&value) {}
    bool has_value() const noexcept { return _value != nullptr; }
    bool empty() const noexcept { return _value == nullptr; }
    explicit operator bool() const noexcept { return _value != nullptr; }
    T* ptr() const noexcept { return _value; }
    T& value() const {
        if (!_value) throw std::bad_optional_access();
        return *_value;
    }
    T* operator->() const { return &value(); }
    T& operator*() const { return value(); }
    void reset() noexcept { _value = nullptr; }
    bool operator==(const OptionalRef& other) const noexcept { return _value == other._value; }
    bool operator!=(const OptionalRef& other) const noexcept { return _value != other._value; }
    bool operator==(std::nullptr_t) const noexcept { return _value == nullptr; }
    bool operator!=(std::nullptr_t) const noexcept { return _value != nullptr; }
    bool operator==(std::nullopt_t) const noexcept { return _value == nullptr; }
    bool operator!=(std::nullopt_t) const noexcept { return _value != nullptr; }

private:
    T* _value = nullptr;
};
}  // namespace darmok
#endif
