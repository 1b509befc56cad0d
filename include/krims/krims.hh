// Minimal krims surface used by gscf - part of the gscf-b200 C++ host layer.
// Not a re-implementation of krims: only what src/gscf/*.hh and a Fock-like plugin touch.
#pragma once
#include <algorithm>
#include <cstddef>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <typeindex>
#include <utility>
#include <vector>

namespace krims {

// ---- exceptions / assertion macros --------------------------------------------------------------
struct ExceptionBase : public std::runtime_error {
  explicit ExceptionBase(const std::string& m) : std::runtime_error(m) {}
};
struct ExcInvalidState : public ExceptionBase {
  explicit ExcInvalidState(const std::string& m) : ExceptionBase("Invalid state: " + m) {}
};
struct ExcIO : public ExceptionBase {
  ExcIO() : ExceptionBase("I/O error") {}
};
struct ExcNotImplemented : public ExceptionBase {
  ExcNotImplemented() : ExceptionBase("not implemented") {}
};
struct ExcInternalError : public ExceptionBase {
  ExcInternalError() : ExceptionBase("internal error") {}
};
struct ExcSizeMismatch : public ExceptionBase {
  ExcSizeMismatch(size_t a, size_t b)
        : ExceptionBase("size mismatch: " + std::to_string(a) + " vs " + std::to_string(b)) {}
};

}  // namespace krims

#ifndef NDEBUG
#define assert_dbg(cond, exc) \
  do {                        \
    if (!(cond)) throw exc;   \
  } while (0)
#else
#define assert_dbg(cond, exc) \
  do {                        \
  } while (0)
#endif
#define assert_throw(cond, exc) \
  do {                          \
    if (!(cond)) throw exc;     \
  } while (0)
#define assert_internal(cond) assert_dbg(cond, ::krims::ExcInternalError())
#define assert_implemented(cond) assert_throw(cond, ::krims::ExcNotImplemented())
#define assert_size(a, b) assert_dbg((a) == (b), ::krims::ExcSizeMismatch((a), (b)))
#define assert_equal(a, b) assert_dbg((a) == (b), ::krims::ExcInternalError())
#define assert_greater(a, b) assert_dbg((a) < (b), ::krims::ExcInternalError())
#define assert_greater_equal(a, b) assert_dbg((a) <= (b), ::krims::ExcInternalError())
#define assert_finite(x) assert_dbg(std::isfinite(x), ::krims::ExcInternalError())

namespace krims {

template <typename... Ts>
using VoidType = void;
template <bool B, typename T = void>
using enable_if_t = typename std::enable_if<B, T>::type;
template <typename T>
using remove_reference_t = typename std::remove_reference<T>::type;

// ---- Range ------------------------------------------------------------------------------------------
template <typename T>
class Range {
 public:
  Range() : m_lo(0), m_hi(0) {}
  Range(T lo, T hi) : m_lo(lo), m_hi(hi) {}
  T lower_bound() const { return m_lo; }
  T upper_bound() const { return m_hi; }
  T length() const { return m_hi > m_lo ? m_hi - m_lo : 0; }
  bool empty() const { return length() == 0; }
  bool operator==(const Range& o) const { return m_lo == o.m_lo && m_hi == o.m_hi; }
  bool operator!=(const Range& o) const { return !(*this == o); }

 private:
  T m_lo, m_hi;
};
template <typename T>
Range<T> range(T n) {
  return Range<T>(0, n);
}
template <typename T>
Range<T> range(T lo, T hi) {
  return Range<T>(lo, hi);
}

// ---- Subscribable -----------------------------------------------------------------------------------
class Subscribable {
 public:
  virtual ~Subscribable() = default;
};
template <typename T>
class SubscriptionPointer {
 public:
  SubscriptionPointer() : m_ptr(nullptr) {}
  SubscriptionPointer(const std::string&, T& obj) : m_ptr(&obj) {}
  T& operator*() const { return *m_ptr; }
  T* operator->() const { return m_ptr; }
  T* get() const { return m_ptr; }
  explicit operator bool() const { return m_ptr != nullptr; }

 private:
  T* m_ptr;  // non-owning: the subscribed object must outlive the pointer (ScfStateBase.hh:175,238)
};
template <typename T>
SubscriptionPointer<T> make_subscription(T& obj, const std::string& who) {
  return SubscriptionPointer<T>(who, obj);
}

// ---- CircularBuffer ---------------------------------------------------------------------------------
template <typename T>
class CircularBuffer {
 public:
  typedef typename std::vector<T>::iterator iterator;
  typedef typename std::vector<T>::const_iterator const_iterator;
  explicit CircularBuffer(size_t max_size = 0) : m_max(max_size) {}
  size_t max_size() const { return m_max; }
  void max_size(size_t m) {
    m_max = m;
    while (m_data.size() > m_max) m_data.erase(m_data.begin());
  }
  size_t size() const { return m_data.size(); }
  bool empty() const { return m_data.empty(); }
  // push_back evicts front() when full (PulayDiisScf.hh:315,318,462)
  void push_back(T v) {
    if (m_max == 0) return;
    if (m_data.size() == m_max) m_data.erase(m_data.begin());
    m_data.push_back(std::move(v));
  }
  void clear() { m_data.clear(); }
  T& front() { return m_data.front(); }
  const T& front() const { return m_data.front(); }
  T& back() { return m_data.back(); }
  const T& back() const { return m_data.back(); }
  iterator begin() { return m_data.begin(); }
  iterator end() { return m_data.end(); }
  const_iterator begin() const { return m_data.begin(); }
  const_iterator end() const { return m_data.end(); }

 private:
  size_t m_max;
  std::vector<T> m_data;  // oldest -> newest
};

// ---- GenMap -----------------------------------------------------------------------------------------
class GenMap {
  struct Entry {
    std::shared_ptr<const void> ptr;
    std::type_index type;
    Entry() : type(typeid(void)) {}
    Entry(std::shared_ptr<const void> p, std::type_index t) : ptr(std::move(p)), type(t) {}
  };

 public:
  struct Item {
    std::string key;
    Entry entry;
    template <typename T, typename = enable_if_t<!std::is_same<typename std::decay<T>::type, GenMap>::value>>
    Item(std::string k, T&& v) : key(std::move(k)) {
      typedef typename std::decay<T>::type V;
      entry = make_entry(std::forward<T>(v), static_cast<V*>(nullptr));
    }
    Item(std::string k, const char* v) : key(std::move(k)) {
      entry = Entry(std::make_shared<const std::string>(v), typeid(std::string));
    }
    Item(std::string k, const GenMap& sub) : key(std::move(k)), is_submap(true), submap(new GenMap(sub)) {}
    bool is_submap = false;
    std::shared_ptr<GenMap> submap;

   private:
    template <typename T, typename V>
    static Entry make_entry(T&& v, V*) {
      return Entry(std::make_shared<const V>(std::forward<T>(v)), typeid(V));
    }
    // shared pointers are stored as pointers to the pointee (at<T> returns const T&)
    template <typename T, typename U>
    static Entry make_entry(T&& v, std::shared_ptr<U>*) {
      return Entry(std::static_pointer_cast<const void>(std::shared_ptr<const U>(v)),
                   typeid(typename std::remove_const<U>::type));
    }
  };

  GenMap() {}
  GenMap(std::initializer_list<Item> il) {
    for (const auto& it : il) insert(it);
  }

  bool exists(const std::string& key) const { return m_map.count(norm(key)) > 0; }

  template <typename T>
  const T& at(const std::string& key) const {
    auto it = m_map.find(norm(key));
    if (it == m_map.end()) throw ExcInvalidState("GenMap: unknown key " + key);
    check_type<T>(it->second, key);
    return *static_cast<const T*>(it->second.ptr.get());
  }
  template <typename T>
  T at(const std::string& key, const T& def) const {
    auto it = m_map.find(norm(key));
    if (it == m_map.end()) return def;
    return convert<T>(it->second, key);
  }
  std::string at(const std::string& key, const char* def) const { return at<std::string>(key, std::string(def)); }

  template <typename T, typename = enable_if_t<!std::is_same<typename std::decay<T>::type, GenMap>::value>>
  void update(const std::string& key, T&& v) {
    insert(Item(key, std::forward<T>(v)));
  }
  void update(const std::string& key, const GenMap& sub) { insert(Item(key, sub)); }
  void update(std::initializer_list<Item> il) {
    for (const auto& it : il) insert(it);
  }

  GenMap submap(const std::string& prefix) const {
    GenMap out;
    const std::string pre = norm(prefix) + "/";
    for (const auto& kv : m_map)
      if (kv.first.compare(0, pre.size(), pre) == 0) out.m_map["/" + kv.first.substr(pre.size())] = kv.second;
    return out;
  }
  size_t size() const { return m_map.size(); }

 private:
  static std::string norm(const std::string& k) {
    std::string s = k;
    while (!s.empty() && s.back() == '/') s.pop_back();
    if (s.empty() || s[0] != '/') s = "/" + s;
    return s;
  }
  void insert(const Item& it) {
    if (it.is_submap) {
      for (const auto& kv : it.submap->m_map) m_map[norm(it.key) + kv.first] = kv.second;
    } else {
      m_map[norm(it.key)] = it.entry;
    }
  }
  template <typename T>
  static void check_type(const Entry& e, const std::string& key) {
    if (e.type != std::type_index(typeid(T)))
      throw ExcInvalidState("GenMap: type mismatch for key " + key);
  }
  // numeric values are convertible between the arithmetic types the reference uses
  template <typename T>
  static enable_if_t<std::is_arithmetic<T>::value, T> convert(const Entry& e, const std::string& key) {
    const void* p = e.ptr.get();
    if (e.type == typeid(T)) return *static_cast<const T*>(p);
    if (e.type == typeid(double)) return static_cast<T>(*static_cast<const double*>(p));
    if (e.type == typeid(float)) return static_cast<T>(*static_cast<const float*>(p));
    if (e.type == typeid(int)) return static_cast<T>(*static_cast<const int*>(p));
    if (e.type == typeid(long)) return static_cast<T>(*static_cast<const long*>(p));
    if (e.type == typeid(unsigned)) return static_cast<T>(*static_cast<const unsigned*>(p));
    if (e.type == typeid(unsigned long)) return static_cast<T>(*static_cast<const unsigned long*>(p));
    if (e.type == typeid(unsigned long long)) return static_cast<T>(*static_cast<const unsigned long long*>(p));
    if (e.type == typeid(long long)) return static_cast<T>(*static_cast<const long long*>(p));
    if (e.type == typeid(bool)) return static_cast<T>(*static_cast<const bool*>(p));
    throw ExcInvalidState("GenMap: type mismatch for key " + key);
  }
  template <typename T>
  static enable_if_t<!std::is_arithmetic<T>::value, T> convert(const Entry& e, const std::string& key) {
    check_type<T>(e, key);
    return *static_cast<const T*>(e.ptr.get());
  }
  std::map<std::string, Entry> m_map;
};

}  // namespace krims
