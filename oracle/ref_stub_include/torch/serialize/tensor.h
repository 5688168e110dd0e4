// Stub for the one torch header the reference's pointnet2 kernel files include (interpolate_gpu.h:4): the .cu files only
// DECLARE at::Tensor wrappers and never use the type, so an incomplete class is all they need.  Lets
// /root/reference/det3d/ops/pointnet2_batch/src/interpolate_gpu.cu compile with plain nvcc, unmodified (oracle/Makefile).
#pragma once
namespace at { class Tensor; }
