// tests/adapter/adapter_bake.cpp — TEST INFRASTRUCTURE: drives vfd::DFSPHSimulation (the reference's class, its header unchanged,
// its body replaced by tests/adapter/DFSPHSimulator.cpp over libvfd_dfsph.so) the way the editor's "Bake" button does
// (Editor/Panels/ComponentPanel.cpp:608-666): description -> fluid objects -> rigid bodies (the reference's own RigidBody:
// SDF + volume map on the host) -> Simulate() -> frames out of the reference's DFSPHParticleBuffer.
//
//   adapter_bake <side> <frames> <frame_length> <out.bin>
// scene: side^3 lattice block (spacing 2r, jittered) in an inverted box, DFSPH with two pinned Jacobi iterations each,
// viscosity and surface tension off — tests/test_gpu_adapter.py builds the same scene for the oracle.
// out.bin: u32 frames, u32 particles, then per frame: f32 MaxVelocityMagnitude, f32 CurrentTimeStep, particles x 9 f32.
#include "pch.h"
#include "Simulation/DFSPH/DFSPHSimulator.h"
#include <cstdio>
#include <cstdlib>

using namespace vfd;

int main(int argc, char** argv)
{
	if (argc < 5) { fprintf(stderr, "usage: adapter_bake <side> <frames> <frame_length> <out.bin> [positions.bin]\n"); return 2; }
	const int side = atoi(argv[1]);
	const unsigned int frames = (unsigned int)atoi(argv[2]);
	const float frameLength = (float)atof(argv[3]);
	const float r = 0.025f, d = 0.05f;

	DFSPHSimulationDescription desc;
	desc.FrameCount = frames;
	desc.FrameLength = frameLength;
	desc.MinPressureSolverIterations = 2; desc.MaxPressureSolverIterations = 2;
	desc.MinDivergenceSolverIterations = 2; desc.MaxDivergenceSolverIterations = 2;
	desc.EnableViscositySolver = false;
	desc.EnableSurfaceTensionSolver = false;
	desc.ParticleRadius = r;

	// positions come from the test (same bytes as the oracle gets)
	std::vector<glm::vec3> positions;
	if (argc > 5) {
		FILE* f = fopen(argv[5], "rb");
		if (!f) { fprintf(stderr, "cannot read %s\n", argv[5]); return 2; }
		glm::vec3 p;
		while (fread(&p, sizeof p, 1, f) == 1) positions.push_back(p);
		fclose(f);
	}
	const float clearance = 4.0f;
	const glm::vec3 boxMax((side + 2 * clearance + 8) * d, (side + 2 * clearance + 4) * d, (side + 2 * clearance) * d);

	Ref<DFSPHSimulation> sim = Ref<DFSPHSimulation>::Create(desc);
	std::vector<Ref<FluidObject>> fluids = { Ref<FluidObject>::Create(positions, glm::vec3(0.0f)) };
	sim->SetFluidObjects(fluids);

	RigidBodyDescription rd;
	rd.Inverted = true;
	rd.Padding = 0.0f;
	rd.CollisionMapResolution = { 10u, 10u, 10u };
	rd.Transform = glm::mat4(1.0f);
	rd.Mesh = Ref<TriangleMesh>::Create(AABB(glm::vec3(0.0f), boxMax.x, boxMax.y, boxMax.z));
	std::vector<Ref<RigidBody>> bodies = { Ref<RigidBody>::Create(rd, sim->GetInfo(), sim->GetKernel()) };
	sim->SetRigidBodies(bodies);

	sim->Simulate();
	if (sim->GetSimulationState() != DFSPHImplementation::SimulationState::Ready) { fprintf(stderr, "bake did not finish\n"); return 1; }

	const unsigned int n = sim->GetParticleCount();
	FILE* out = fopen(argv[4], "wb");
	if (!out) { fprintf(stderr, "cannot write %s\n", argv[4]); return 2; }
	fwrite(&frames, 4, 1, out); fwrite(&n, 4, 1, out);
	Ref<DFSPHParticleBuffer> buffer = sim->GetParticleFrameBuffer();
	for (unsigned int i = 0; i < frames; i++)
	{
		const Ref<DFSPHParticleFrame>& frame = buffer->GetFrame(i);
		fwrite(&frame->MaxVelocityMagnitude, 4, 1, out);
		fwrite(&frame->CurrentTimeStep, 4, 1, out);
		fwrite(frame->ParticleData.data(), sizeof(DFSPHParticleSimple), n, out);
	}
	fclose(out);
	const DFSPHDebugInfo& dbg = sim->GetDebugInfo();
	printf("adapter bake: %u particles, %u frames, %u steps, rigid bodies %u, dt %.6g\n", n, frames, dbg.IterationCount, sim->GetRigidBodyCount(), sim->GetCurrentTimeStepSize());
	return 0;
}
