// vsf_sequence_driver — the frame-source side of the drop-in (reference:
// src/slam_frontend_main.cc:236-328 reads a rosbag, feeds ObserveOdometry / ObserveImage and
// writes ONE SLAMProblem message, :369-374).  Here the frames come from the synthetic stereo
// source, the sequence is sharded into contiguous pose ranges over the ranks of one node (one
// process per GPU, no data-path collective), and the ranks' pieces of the message are gathered
// to rank 0 over NCCL (libvsf_nccl.so) and written in ROS1 wire format.
//
//   vsf_sequence_driver --poses 400 --features 2000 --window 10 --desc-bytes 61 --out problem.bin
//   RANK / WORLD_SIZE / LOCAL_RANK (torchrun-style) or --rank / --world / --device select the shard;
//   --rendezvous FILE is where rank 0 leaves the NCCL unique id for the others.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "synthetic_source.h"
#include "vsf_nccl.h"

namespace {
const char* Env(const char* k, const char* d) {
  const char* v = std::getenv(k);
  return v ? v : d;
}
}  // namespace

int main(int argc, char** argv) {
  int poses = 200, features = 2000, window = 10, desc_bytes = 61, in_flight = 3, exact = 1;
  int world = std::atoi(Env("WORLD_SIZE", "1")), rank = std::atoi(Env("RANK", "0"));
  int device = std::atoi(Env("LOCAL_RANK", "0"));
  unsigned long long seed = 1;
  std::string out_path = "slam_problem.bin", rendezvous = "/tmp/vsf_nccl_id";
  for (int i = 1; i + 1 < argc; i += 2) {
    const std::string k = argv[i];
    const char* v = argv[i + 1];
    if (k == "--poses") poses = std::atoi(v);
    else if (k == "--features") features = std::atoi(v);
    else if (k == "--window") window = std::atoi(v);
    else if (k == "--desc-bytes") desc_bytes = std::atoi(v);
    else if (k == "--in-flight") in_flight = std::atoi(v);
    else if (k == "--exact-sort") exact = std::atoi(v);
    else if (k == "--world") world = std::atoi(v);
    else if (k == "--rank") rank = std::atoi(v);
    else if (k == "--device") device = std::atoi(v);
    else if (k == "--seed") seed = std::strtoull(v, nullptr, 10);
    else if (k == "--out") out_path = v;
    else if (k == "--rendezvous") rendezvous = v;
    else {
      std::fprintf(stderr, "unknown option %s\n", k.c_str());
      return 2;
    }
  }
  try {
    const slam::FrontendConfig rig = slam::SyntheticRig(device, features, desc_bytes, window, exact != 0);
    slam::SyntheticStereoConfig sc;
    sc.features = features;
    sc.landmark_stride = std::max(1, features / 10);
    sc.seed = seed;
    slam::SyntheticStereoSource source(rig, sc);
    const unsigned long long base = (unsigned long long)poses / world, rem = (unsigned long long)poses % world;
    const unsigned long long first = rank * base + std::min<unsigned long long>(rank, rem);
    const unsigned long long last = first + base + ((unsigned long long)rank < rem ? 1 : 0);

    vsf_nccl_comm* comm = nullptr;
    if (world > 1) {
      char id[VSF_NCCL_UNIQUE_ID_BYTES];
      if (rank == 0) {
        if (vsf_nccl_unique_id(id)) throw std::runtime_error("ncclGetUniqueId failed");
        std::ofstream f(rendezvous + ".tmp", std::ios::binary);
        f.write(id, sizeof(id));
        f.close();
        std::rename((rendezvous + ".tmp").c_str(), rendezvous.c_str());
      } else {
        for (int tries = 0;; ++tries) {
          std::ifstream f(rendezvous, std::ios::binary);
          if (f && f.read(id, sizeof(id))) break;
          if (tries > 6000) throw std::runtime_error("no NCCL id at " + rendezvous);
          std::this_thread::sleep_for(std::chrono::milliseconds(10));
        }
      }
      if (vsf_nccl_comm_create(id, world, rank, device, &comm)) throw std::runtime_error("ncclCommInitRank failed");
    }

    const auto t0 = std::chrono::steady_clock::now();
    double host_us[3] = {0, 0, 0};
    const slam::SLAMProblemPiece piece = slam::RunSequenceShard(rig, source, first, last, in_flight, host_us);
    const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    std::vector<slam::SLAMProblemPiece> pieces;
    if (world > 1) {
      const std::vector<uint8_t> blob = piece.Pack();
      std::vector<size_t> sizes(world);
      void* all = nullptr;
      if (vsf_nccl_gather_bytes(comm, blob.data(), blob.size(), 0, &all, sizes.data()))
        throw std::runtime_error(std::string("gather failed: ") + vsf_nccl_last_error(comm));
      if (rank == 0) {
        size_t off = 0;
        for (int r = 0; r < world; ++r) {
          pieces.push_back(slam::SLAMProblemPiece::Unpack(static_cast<const uint8_t*>(all) + off, sizes[r]));
          off += sizes[r];
        }
        std::free(all);
      }
      vsf_nccl_comm_destroy(comm);
    } else {
      pieces.push_back(piece);
    }
    std::fprintf(stderr, "rank %d: poses [%llu, %llu) in %.3f s incl. context creation, %u nodes, %u vision factors; host per frame: "
                 "source %.0f us, SubmitFeatures %.0f us, CollectFeatures %.0f us\n",
                 rank, first, last, el, piece.n_nodes, piece.n_vision_factors, host_us[0], host_us[1], host_us[2]);
    if (rank == 0) {
      const std::vector<uint8_t> wire = slam::MergeSLAMProblemPieces(pieces);
      std::ofstream f(out_path, std::ios::binary);
      f.write(reinterpret_cast<const char*>(wire.data()), std::streamsize(wire.size()));
      std::printf("{\"poses\": %d, \"world\": %d, \"features\": %d, \"window\": %d, \"desc_bytes\": %d, "
                  "\"bytes\": %zu, \"rank0_seconds\": %.4f, \"out\": \"%s\"}\n",
                  poses, world, features, window, desc_bytes, wire.size(), el, out_path.c_str());
      if (world > 1) std::remove(rendezvous.c_str());
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "vsf_sequence_driver (rank %d): %s\n", rank, e.what());
    return 1;
  }
  return 0;
}
