// oracle/shim/pch.h — TEST INFRASTRUCTURE (oracle build only, never shipped).
// Replacement for the reference's Windows-only precompiled header
// (/root/reference/VFD/Source/pch.h:1-58) so that the reference's own DFSPH
// solver sources compile headless on Linux.  Only what the solver path needs.
#ifndef VFD_ORACLE_SHIM_PCH_H
#define VFD_ORACLE_SHIM_PCH_H

#include <iostream>
#include <sstream>
#include <fstream>
#include <memory>
#include <unordered_set>
#include <unordered_map>
#include <mutex>
#include <functional>
#include <future>
#include <vector>
#include <string>
#include <map>
#include <algorithm>
#include <random>
#include <queue>
#include <list>
#include <set>
#include <stack>
#include <deque>
#include <thread>
#include <filesystem>
#include <immintrin.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cassert>
#include <array>
#include <numeric>
#include <atomic>
#include <chrono>
#include <cstddef>
#include <cstdint>
#include <iterator>
#include <limits>

#include <cuda.h>
#include <cuda_runtime.h>

#define GLM_ENABLE_EXPERIMENTAL
#ifdef __CUDACC__
#define GLM_FORCE_CUDA
#endif
#include <glm/glm.hpp>
#include <glm/ext.hpp>
#include <glm/gtc/type_ptr.hpp>
#include <glm/gtx/quaternion.hpp>
#include <glm/gtc/matrix_access.hpp>
#include <glm/gtx/norm.hpp>
#include <glm/gtx/component_wise.hpp>
#include <glm/gtx/matrix_decompose.hpp>

// Debug.h replacements (reference: VFD/Source/Debug/Debug.h:90-125)
#define LOG(...)
#define WARN(...)
#define ERR(...)
#define ASSERT(...)
#define COMPUTE_SAFE(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %d at %s:%d\n", (int)e__, __FILE__, __LINE__); abort(); } } while (0)
#define COMPUTE_CHECK(msg)

// Intrusive ref counting with the same surface as VFD/Source/Core/Ref.h:5-120
namespace vfd {
    class RefCounted {
    public:
        virtual ~RefCounted() = default;
        void IncRefCount() const { ++m_RefCount; }
        void DecRefCount() const { --m_RefCount; }
        uint32_t GetRefCount() const { return m_RefCount.load(); }
    private:
        mutable std::atomic<uint32_t> m_RefCount{0};
    };

    template<typename T>
    class Ref {
    public:
        Ref() : m_Instance(nullptr) {}
        Ref(std::nullptr_t) : m_Instance(nullptr) {}
        Ref(T* instance) : m_Instance(instance) { IncRef(); }
        Ref(const Ref<T>& other) : m_Instance(other.m_Instance) { IncRef(); }
        template<typename T2> Ref(const Ref<T2>& other) : m_Instance(static_cast<T*>(other.m_Instance)) { IncRef(); }
        ~Ref() { DecRef(); }
        Ref& operator=(std::nullptr_t) { DecRef(); m_Instance = nullptr; return *this; }
        Ref& operator=(const Ref<T>& other) { other.IncRef(); DecRef(); m_Instance = other.m_Instance; return *this; }
        template<typename T2> Ref& operator=(const Ref<T2>& other) { other.IncRef(); DecRef(); m_Instance = other.m_Instance; return *this; }
        operator bool() { return m_Instance != nullptr; }
        operator bool() const { return m_Instance != nullptr; }
        T* operator->() { return m_Instance; }
        const T* operator->() const { return m_Instance; }
        T& operator*() { return *m_Instance; }
        const T& operator*() const { return *m_Instance; }
        T* Raw() { return m_Instance; }
        const T* Raw() const { return m_Instance; }
        template<typename... Args> static Ref<T> Create(Args&&... args) { return Ref<T>(new T(std::forward<Args>(args)...)); }
        bool operator==(const Ref<T>& o) const { return m_Instance == o.m_Instance; }
        bool operator!=(const Ref<T>& o) const { return m_Instance != o.m_Instance; }
    private:
        void IncRef() const { if (m_Instance) m_Instance->IncRefCount(); }
        void DecRef() const {
            if (m_Instance) { m_Instance->DecRefCount();
                if (m_Instance->GetRefCount() == 0) { delete m_Instance; m_Instance = nullptr; } }
        }
        template<class T2> friend class Ref;
        mutable T* m_Instance;
    };
}
// MSVC-ism: the reference is written against MSVC's RAND_MAX = 32767.  BoundingSphere.h:154 forms
// i * rand() in int before dividing by RAND_MAX; with glibc's RAND_MAX = 2^31 - 1 the product overflows, the "random
// permutation" index goes negative and the smallest-enclosing-sphere code reads and writes outside its vertex array —
// after which MeshDistance's sphere tree prunes the wrong branches now and then (observed: 193 078 of 200 000 queries of
// one call wrong, the repeated call right).  Give the reference the 15-bit rand() it was written for.
#include <cstdlib>
static inline int vfd_ref_rand15() { return std::rand() & 0x7fff; }
#undef RAND_MAX
#define RAND_MAX 0x7fff
#define rand vfd_ref_rand15
#endif
