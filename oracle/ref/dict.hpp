// TEST INFRASTRUCTURE — not part of the product path.
//
// Minimal stand-in for the cppdict header (`dict.hpp`) that the reference pulls at configure time
// from LaboratoryOfPlasmaPhysics/cppdict (res/cmake/dep/cppdict.cmake:3-9, branch master, not
// vendored under /root/reference).  Only the API surface the hot path touches is provided:
//   cppdict::Dict<Ts...>  operator[] (creating / const-throwing), operator=(T), to<T>(),
//   to<T>(default), contains(), visit(fn(key,value)), cppdict::get_value(dict, "a/b", default)
// Users in the reference: src/initializer/data_provider.hpp:52-55, src/core/errors.hpp:23-46,
// src/core/numerics/ohm/ohm.hpp:27, particle_initializer_factory.hpp:50.
// It is a configuration tree only; no arithmetic of the path lives in it.
#ifndef PHB_ORACLE_DICT_SHIM_HPP
#define PHB_ORACLE_DICT_SHIM_HPP

#include <map>
#include <memory>
#include <string>
#include <variant>
#include <stdexcept>
#include <type_traits>

namespace cppdict
{
template<typename... Types>
struct Dict
{
    using This     = Dict<Types...>;
    using children = std::map<std::string, std::shared_ptr<This>>;
    using data_t   = std::variant<std::monostate, children, Types...>;

    data_t data{};

    Dict() = default;
    Dict(Dict const& other) { *this = other; }
    Dict(Dict&&) = default;
    Dict& operator=(Dict&&) = default;

    // deep copy so that sub-dicts do not alias
    Dict& operator=(Dict const& other)
    {
        if (this == &other)
            return *this;
        if (std::holds_alternative<children>(other.data))
        {
            children mine;
            for (auto const& [k, v] : std::get<children>(other.data))
                mine[k] = std::make_shared<This>(*v);
            data = std::move(mine);
        }
        else
            data = other.data;
        return *this;
    }

    template<typename T, typename = std::enable_if_t<!std::is_same_v<std::decay_t<T>, This>>>
    Dict& operator=(T&& value)
    {
        using V = std::decay_t<T>;
        if constexpr ((std::is_same_v<V, Types> || ...))
            data = std::forward<T>(value);
        else if constexpr (std::is_convertible_v<V, std::string>
                           && (std::is_same_v<std::string, Types> || ...))
            data = std::string{value};
        else
            static_assert(sizeof(T) == 0, "type not storable in this Dict");
        return *this;
    }

    bool isEmpty() const { return std::holds_alternative<std::monostate>(data); }
    bool isNode() const { return std::holds_alternative<children>(data); }
    bool isValue() const { return !isEmpty() && !isNode(); }

    This& operator[](std::string const& key)
    {
        if (isEmpty())
            data = children{};
        if (!isNode())
            throw std::runtime_error("cppdict shim: operator[] on a leaf: " + key);
        auto& kids = std::get<children>(data);
        auto it    = kids.find(key);
        if (it == kids.end())
            it = kids.emplace(key, std::make_shared<This>()).first;
        return *it->second;
    }

    This const& operator[](std::string const& key) const
    {
        if (!isNode())
            throw std::runtime_error("cppdict shim: const operator[] on non-node: " + key);
        auto const& kids = std::get<children>(data);
        auto it          = kids.find(key);
        if (it == kids.end())
            throw std::runtime_error("cppdict shim: missing key: " + key);
        return *it->second;
    }

    bool contains(std::string const& key) const
    {
        return isNode() && std::get<children>(data).count(key) > 0;
    }

    std::size_t size() const { return isNode() ? std::get<children>(data).size() : 0; }

    template<typename T>
    T& to()
    {
        if (auto* p = std::get_if<T>(&data))
            return *p;
        throw std::runtime_error("cppdict shim: to<T>() wrong type or empty");
    }
    template<typename T>
    T const& to() const
    {
        if (auto const* p = std::get_if<T>(&data))
            return *p;
        throw std::runtime_error("cppdict shim: to<T>() wrong type or empty");
    }
    template<typename T>
    T to(T const& dflt) const
    {
        if (auto const* p = std::get_if<T>(&data))
            return *p;
        return dflt;
    }

    // visits the leaf values directly under this node
    template<typename Fn>
    void visit(Fn&& fn) const
    {
        if (!isNode())
            return;
        for (auto const& [k, v] : std::get<children>(data))
        {
            std::visit(
                [&](auto const& val) {
                    using V = std::decay_t<decltype(val)>;
                    if constexpr (!std::is_same_v<V, std::monostate> && !std::is_same_v<V, children>)
                        fn(k, val);
                },
                v->data);
        }
    }
};


template<typename T, typename... Ts>
T get_value(Dict<Ts...> const& dict, std::string path, T const& dflt)
{
    Dict<Ts...> const* node = &dict;
    std::size_t pos         = 0;
    while (true)
    {
        auto const next = path.find('/', pos);
        auto const key  = path.substr(pos, next == std::string::npos ? next : next - pos);
        if (!node->contains(key))
            return dflt;
        node = &(*node)[key];
        if (next == std::string::npos)
            break;
        pos = next + 1;
    }
    return node->template to<T>(dflt);
}

template<typename T, typename... Ts>
void add(std::string path, T&& value, Dict<Ts...>& dict)
{
    Dict<Ts...>* node = &dict;
    std::size_t pos   = 0;
    while (true)
    {
        auto const next = path.find('/', pos);
        auto const key  = path.substr(pos, next == std::string::npos ? next : next - pos);
        node            = &(*node)[key];
        if (next == std::string::npos)
            break;
        pos = next + 1;
    }
    *node = std::forward<T>(value);
}

} // namespace cppdict

#endif
