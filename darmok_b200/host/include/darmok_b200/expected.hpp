// darmok_b200/expected.hpp -- darmok::expected / darmok::unexpected for the shim.
//
// darmok is C++20 (CMakeLists.txt:28) and aliases tl::expected (include/darmok/expected.hpp:3-12).  Inside a darmok build that
// header is on the include path and the shim uses the very same alias; a stand-alone C++23 build uses std::expected; a
// stand-alone C++20 build without tl gets the small class below (the subset the shim uses: value / error access, bool
// conversion, construction from a value, from `unexpected`, and `{}` for expected<void, E>).
#pragma once

#if __has_include(<tl/expected.hpp>)
#include <tl/expected.hpp>
namespace darmok {
template <class T, class E>
using expected = tl::expected<T, E>;
template <class E>
using unexpected = tl::unexpected<E>;
}  // namespace darmok
#elif defined(__cpp_lib_expected) || (__cplusplus >= 202302L && __has_include(<expected>))
#include <expected>
namespace darmok {
template <class T, class E>
using expected = std::expected<T, E>;
template <class E>
using unexpected = std::unexpected<E>;
}  // namespace darmok
#else
#include <optional>
#include <type_traits>
#include <utility>
namespace darmok {
template <class E>
class unexpected {
public:
    explicit unexpected(E e) : _error(std::move(e)) {}
    const E& error() const& noexcept { return _error; }
    E& error() & noexcept { return _error; }
    E&& error() && noexcept { return std::move(_error); }

private:
    E _error;
};
template <class E>
unexpected(E) -> unexpected<E>;

template <class T, class E>
class expected {
public:
    expected() : _value(T{}) {}
    expected(T v) : _value(std::move(v)) {}
    template <class U, std::enable_if_t<std::is_convertible_v<U, T> && !std::is_same_v<std::decay_t<U>, T> && !std::is_same_v<std::decay_t<U>, expected>, int> = 0>
    expected(U&& v) : _value(T(std::forward<U>(v))) {}
    template <class G>
    expected(unexpected<G> u) : _error(E(std::move(u).error())) {}
    bool has_value() const noexcept { return _value.has_value(); }
    explicit operator bool() const noexcept { return has_value(); }
    T& value() & { return *_value; }
    const T& value() const& { return *_value; }
    T&& value() && { return std::move(*_value); }
    T& operator*() & noexcept { return *_value; }
    const T& operator*() const& noexcept { return *_value; }
    T&& operator*() && noexcept { return std::move(*_value); }
    T* operator->() noexcept { return &*_value; }
    const T* operator->() const noexcept { return &*_value; }
    E& error() & noexcept { return *_error; }
    const E& error() const& noexcept { return *_error; }
    E&& error() && noexcept { return std::move(*_error); }

private:
    std::optional<T> _value;
    std::optional<E> _error;
};
template <class E>
class expected<void, E> {
public:
    expected() = default;
    template <class G>
    expected(unexpected<G> u) : _error(E(std::move(u).error())) {}
    bool has_value() const noexcept { return !_error.has_value(); }
    explicit operator bool() const noexcept { return has_value(); }
    void value() const noexcept {}
    E& error() & noexcept { return *_error; }
    const E& error() const& noexcept { return *_error; }
    E&& error() && noexcept { return std::move(*_error); }

private:
    std::optional<E> _error;
};
}  // namespace darmok
#endif
