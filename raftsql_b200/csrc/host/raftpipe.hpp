// raftpipe.hpp — the drop-in seam in C++, in the reference's shape (reference raftpipe.go:3-17):
//
//   auto rp = NewRaftPipe(id, peers, proposeC);      // raftpipe.go:9-12
//   rp->ProposeC (send strings)   rp->CommitC (committed strings; nullptr = "log replayed")   rp->ErrorC
//   rp->Close();                                     // raftpipe.go:14-17: close(ProposeC); return <-ErrorC
//
// Channel protocol preserved from the reference: every replayed entry, then one nil, then live
// entries (raft.go:57-61,130-132); log order, empty / conf-change entries skipped (raft.go:84-86); shutdown =
// caller closes ProposeC -> CommitC closed, ErrorC closed without a value (raft.go:216-217,241-243,191-196);
// a failure closes CommitC, sends the error on ErrorC, closes ErrorC (writeError, raft.go:136-142).
#pragma once

#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "chan.hpp"
#include "hostnode.hpp"

namespace raftsql {

using StrChan = Chan<std::string>;
using CommitChan = Chan<std::shared_ptr<std::string>>;  // <-chan *string; nullptr is the nil sentinel

struct RaftPipeOptions {
  double tick_seconds = 0.1;                    // the reference's 100 ms ticker (raft.go:207)
  std::string waldir = "auto";                  // "auto": raftsql-<id> (raft.go:69); "": no WAL
  std::shared_ptr<LocalTransport> transport;    // default: one shared in-process transport per peer list
  // consensus core for (npeers, id); default: the GPU engine through the C-ABI (make_engine_core)
  std::function<std::unique_ptr<Core>(uint32_t, uint32_t)> core_factory;
};

class RaftPipe {
 public:
  std::shared_ptr<StrChan> ProposeC;
  std::shared_ptr<CommitChan> CommitC;
  std::shared_ptr<StrChan> ErrorC;
  // reference raftpipe.go:14-17; returns the error text, empty on a clean shutdown
  std::string Close();
  ~RaftPipe();
  HostNode *node() { return node_.get(); }  // introspection (role, term, commit): not part of the seam

 private:
  friend std::unique_ptr<RaftPipe> NewRaftPipe(int, const std::vector<std::string> &, std::shared_ptr<StrChan>,
                                               const RaftPipeOptions &);
  std::shared_ptr<HostNode> node_;
  std::thread thread_;
};

std::unique_ptr<RaftPipe> NewRaftPipe(int id, const std::vector<std::string> &peers, std::shared_ptr<StrChan> proposeC,
                                      const RaftPipeOptions &opt = RaftPipeOptions());

// The product core: one engine (G = 1, R = npeers, self_id = id) on `device` behind include/mrq.h.
std::unique_ptr<Core> make_engine_core(uint32_t npeers, uint32_t id, int device = 0);

std::shared_ptr<LocalTransport> transport_for(const std::vector<std::string> &peers);

// ---- the multi-group seam --------------------------------------------------------------------
// G raft groups of one node behind per-group channels over ONE engine of G groups, ticked once per tick for all:
//   mp->ProposeC[g] / mp->CommitC[g]  — each with the protocol above (replay, nil, live entries of group g in order)
//   mp->ErrorC, mp->Close()           — one per node; Close() closes every ProposeC and returns <-ErrorC
struct MultiRaftPipeOptions {
  double tick_seconds = 0.1;
  std::string waldir = "auto";  // "auto": raftsql-<id> (one group-commit file for all groups, MultiWal); "": no WAL
  std::shared_ptr<MultiLocalTransport> transport;  // required when several nodes share a process
  // consensus core for (npeers, id, n_groups); default: one GPU engine of n_groups groups (make_engine_multicore)
  std::function<std::unique_ptr<MultiCore>(uint32_t, uint32_t, size_t)> core_factory;
};

class MultiRaftPipe {
 public:
  std::vector<std::shared_ptr<StrChan>> ProposeC;
  std::vector<std::shared_ptr<CommitChan>> CommitC;
  std::shared_ptr<StrChan> ErrorC;
  std::string Close();
  ~MultiRaftPipe();
  MultiHostNode *node() { return node_.get(); }

 private:
  friend std::unique_ptr<MultiRaftPipe> NewMultiRaftPipe(int, const std::vector<std::string> &, size_t,
                                                         const MultiRaftPipeOptions &);
  std::shared_ptr<MultiHostNode> node_;
  std::thread thread_;
};

std::unique_ptr<MultiRaftPipe> NewMultiRaftPipe(int id, const std::vector<std::string> &peers, size_t n_groups,
                                                const MultiRaftPipeOptions &opt = MultiRaftPipeOptions());

// The product multi-group core: one engine (G = n_groups, R = npeers, self_id = id) behind include/mrq.h.
std::unique_ptr<MultiCore> make_engine_multicore(uint32_t npeers, uint32_t id, size_t n_groups, int device = 0);

}  // namespace raftsql
