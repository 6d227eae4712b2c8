// oracle/shim — TEST INFRASTRUCTURE: a TriangleMesh that only knows how to be a box
// (same 8 vertices / 12 triangles, in the same order, as the reference's
// TriangleMesh(const AABB&) — VFD/Source/Renderer/Mesh/TriangleMesh.cpp:18-37) or to
// take raw arrays.  Avoids tinyobjloader / <xhash> / OpenGL.
#ifndef VFD_ORACLE_SHIM_TRIANGLE_MESH_H
#define VFD_ORACLE_SHIM_TRIANGLE_MESH_H
#include "pch.h"
#include "Renderer/VertexArray.h"
#include "Core/Structures/AxisAlignedBoundingBox.h"
namespace vfd {
    class TriangleMesh : public RefCounted {
    public:
        TriangleMesh() = default;
        TriangleMesh(const AABB& b) {
            const glm::vec3 p = b.position;
            m_Vertices = {
                {p.x, p.y, p.z}, {p.x + b.width, p.y, p.z}, {p.x + b.width, p.y, p.z + b.depth}, {p.x, p.y, p.z + b.depth},
                {p.x, p.y + b.height, p.z}, {p.x + b.width, p.y + b.height, p.z},
                {p.x + b.width, p.y + b.height, p.z + b.depth}, {p.x, p.y + b.height, p.z + b.depth} };
            m_Triangles = { {0,1,2},{0,2,3},{4,7,6},{4,6,5},{0,3,7},{0,7,4},{1,5,6},{1,6,2},{0,4,5},{0,5,1},{3,2,6},{3,6,7} };
        }
        TriangleMesh(const std::vector<glm::vec3>& v, const std::vector<glm::uvec3>& t) : m_Vertices(v), m_Triangles(t) {}
        const std::vector<glm::vec3>& GetVertices() { return m_Vertices; }
        const std::vector<glm::uvec3>& GetTriangles() { return m_Triangles; }
    private:
        std::vector<glm::vec3> m_Vertices;
        std::vector<glm::uvec3> m_Triangles;
    };
}
#endif
