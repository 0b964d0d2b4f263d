#include "RAJAPerfSuite.hpp"

#include <iostream>

#include "../algorithm/Algorithm.hpp"
#include "../apps/Apps.hpp"
#include "../comm/Comm.hpp"
#include "../basic/Basic.hpp"
#include "../polybench/Polybench.hpp"
#include "../stream/Stream.hpp"

namespace rajaperf {

static const std::string GroupNames[] = {"Basic", "Polybench", "Stream", "Apps", "Algorithm", "Comm", "Unknown Group"};

static const std::string KernelNames[] = {
  "Basic_INDEXLIST", "Basic_INDEXLIST_3LOOP",
  "Polybench_GEMM",
  "Stream_ADD", "Stream_COPY", "Stream_DOT", "Stream_MUL", "Stream_TRIAD",
  "Apps_CONVECTION3DPA", "Apps_DIFFUSION3DPA", "Apps_LTIMES", "Apps_MASS3DPA",
  "Algorithm_SCAN", "Algorithm_SORT", "Algorithm_SORTPAIRS", "Algorithm_REDUCE_SUM", "Algorithm_MEMSET", "Algorithm_MEMCPY",
  "Comm_HALO_PACKING", "Comm_HALO_PACKING_FUSED", "Comm_HALO_SENDRECV", "Comm_HALO_EXCHANGE", "Comm_HALO_EXCHANGE_FUSED",
  "Unknown Kernel"
};

static const std::string VariantNames[] = {"Base_Seq", "Base_OpenMP", "Base_B200", "Unknown Variant"};

const std::string& getGroupName(GroupID gid) { return GroupNames[gid]; }
const std::string& getFullKernelName(KernelID kid) { return KernelNames[kid]; }
std::string getKernelName(KernelID kid)
{
  const std::string& full = KernelNames[kid];
  return full.substr(full.find('_') + 1);
}
GroupID getKernelGroup(KernelID kid)
{
  if (kid <= Basic_INDEXLIST_3LOOP) return Basic;
  if (kid <= Polybench_GEMM) return Polybench;
  if (kid <= Stream_TRIAD) return Stream;
  if (kid <= Apps_MASS3DPA) return Apps;
  if (kid <= Algorithm_MEMCPY) return Algorithm;
  return Comm;
}
const std::string& getVariantName(VariantID vid) { return VariantNames[vid]; }
bool isVariantAvailable(VariantID vid) { return vid == Base_B200; }
bool isVariantGPU(VariantID vid) { return vid == Base_B200; }

KernelBase* getKernelObject(KernelID kid, const RunParams& p)
{
  switch (kid) {
    case Basic_INDEXLIST: return new basic::INDEXLIST(p);
    case Basic_INDEXLIST_3LOOP: return new basic::INDEXLIST_3LOOP(p);
    case Polybench_GEMM: return new polybench::POLYBENCH_GEMM(p);
    case Stream_ADD: return new stream::ADD(p);
    case Stream_COPY: return new stream::COPY(p);
    case Stream_DOT: return new stream::DOT(p);
    case Stream_MUL: return new stream::MUL(p);
    case Stream_TRIAD: return new stream::TRIAD(p);
    case Apps_CONVECTION3DPA: return new apps::CONVECTION3DPA(p);
    case Apps_DIFFUSION3DPA: return new apps::DIFFUSION3DPA(p);
    case Apps_LTIMES: return new apps::LTIMES(p);
    case Apps_MASS3DPA: return new apps::MASS3DPA(p);
    case Algorithm_SCAN: return new algorithm::SCAN(p);
    case Algorithm_SORT: return new algorithm::SORT(p);
    case Algorithm_SORTPAIRS: return new algorithm::SORTPAIRS(p);
    case Algorithm_REDUCE_SUM: return new algorithm::REDUCE_SUM(p);
    case Algorithm_MEMSET: return new algorithm::MEMSET(p);
    case Algorithm_MEMCPY: return new algorithm::MEMCPY(p);
    case Comm_HALO_PACKING: return new comm::HALO_PACKING(p);
    case Comm_HALO_PACKING_FUSED: return new comm::HALO_PACKING_FUSED(p);
    case Comm_HALO_SENDRECV: return new comm::HALO_SENDRECV(p);
    case Comm_HALO_EXCHANGE: return new comm::HALO_EXCHANGE(p);
    case Comm_HALO_EXCHANGE_FUSED: return new comm::HALO_EXCHANGE_FUSED(p);
    default: getCout() << "\n Unknown Kernel ID = " << kid << std::endl; return nullptr;
  }
}

std::ostream& getCout() { return std::cout; }

}  // namespace rajaperf
