// oracle/shim — TEST INFRASTRUCTURE: no-op stand-ins for the OpenGL buffer classes the
// reference's DFSPHParticleBuffer touches (API per VFD/Source/Renderer/Buffers/VertexBuffer.h:25-81,
// VFD/Source/Renderer/VertexArray.h:8-27).  Rendering is out of scope.
#ifndef VFD_ORACLE_SHIM_VERTEX_ARRAY_H
#define VFD_ORACLE_SHIM_VERTEX_ARRAY_H
#include "pch.h"
namespace vfd {
    enum class ShaderDataType { None = 0, Bool, Int, Uint, Float, Float2, Float3, Float4, Mat3, Mat4 };
    struct BufferElement { ShaderDataType Type; std::string Name; BufferElement(ShaderDataType t, const std::string& n) : Type(t), Name(n) {} };
    struct BufferLayout { BufferLayout() = default; BufferLayout(std::initializer_list<BufferElement>) {} };
    class VertexBuffer : public RefCounted {
    public:
        VertexBuffer(uint32_t) {}
        void SetLayout(const BufferLayout&) {}
        void SetData(uint32_t, uint32_t, const void*) {}
        void Bind() const {}
        void Unbind() const {}
    };
    class VertexArray : public RefCounted {
    public:
        void AddVertexBuffer(Ref<VertexBuffer>&) {}
        void Bind() const {}
        void Unbind() const {}
    };
}
#endif
