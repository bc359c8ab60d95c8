// Single-process functional MPI shim (TEST INFRASTRUCTURE ONLY) -- see mpi.h in this directory.
//
// Point-to-point semantics that the reference's halo engine relies on (nix/chunk.hpp:464-543,
// nix/chunk.cpp:288-395):
//   * Isend is buffered: the payload is copied into a FIFO keyed by (communicator, tag) at once.
//   * Irecv matches the head of that FIFO if present, otherwise stays pending until Wait/Test.
//   * Iprobe reports the byte size of the head message; MPI_PROC_NULL peers always "match" with 0 B.
// All calls are serialised by one mutex because the reference calls them from OpenMP workers.
#include "mpi.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

namespace
{
struct Pending {
  void* buf;
  int   maxbytes;
  int   comm;
  int   tag;
  bool  active;
};

using Key = std::pair<int, int>;

std::mutex                                         g_mutex;
std::map<Key, std::deque<std::vector<uint8_t>>>    g_mailbox;
std::vector<Pending>                               g_pending;
std::vector<int>                                   g_free_slots;
std::atomic<int>                                   g_next_comm{1};

// try to complete a pending receive; caller holds the mutex
bool try_complete(int slot)
{
  Pending& p = g_pending[slot];
  if (!p.active)
    return true;
  auto it = g_mailbox.find(Key(p.comm, p.tag));
  if (it == g_mailbox.end() || it->second.empty())
    return false;
  std::vector<uint8_t>& msg = it->second.front();
  if ((int)msg.size() > p.maxbytes) {
    std::fprintf(stderr, "[mpi_shim] message truncated: %zu > %d (comm %d tag %d)\n", msg.size(),
                 p.maxbytes, p.comm, p.tag);
    std::abort();
  }
  if (!msg.empty())
    std::memcpy(p.buf, msg.data(), msg.size());
  it->second.pop_front();
  p.active = false;
  g_free_slots.push_back(slot);
  return true;
}

bool test_one(MPI_Request* req)
{
  if (*req == MPI_REQUEST_NULL)
    return true;
  std::lock_guard<std::mutex> lock(g_mutex);
  if (try_complete(*req)) {
    *req = MPI_REQUEST_NULL;
    return true;
  }
  return false;
}

void wait_one(MPI_Request* req)
{
  auto t0 = std::chrono::steady_clock::now();
  while (!test_one(req)) {
    std::this_thread::yield();
    auto dt = std::chrono::steady_clock::now() - t0;
    if (dt > std::chrono::seconds(60)) {
      std::fprintf(stderr, "[mpi_shim] deadlock: receive never matched\n");
      std::abort();
    }
  }
}
} // namespace

extern "C" {

int MPI_Init(int*, char***)
{
  return MPI_SUCCESS;
}

int MPI_Init_thread(int*, char***, int required, int* provided)
{
  *provided = required;
  return MPI_SUCCESS;
}

int MPI_Finalize(void)
{
  return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm, int code)
{
  std::fprintf(stderr, "[mpi_shim] MPI_Abort(%d)\n", code);
  std::abort();
  return MPI_SUCCESS;
}

int MPI_Comm_rank(MPI_Comm, int* rank)
{
  *rank = 0;
  return MPI_SUCCESS;
}

int MPI_Comm_size(MPI_Comm, int* size)
{
  *size = 1;
  return MPI_SUCCESS;
}

int MPI_Comm_dup(MPI_Comm, MPI_Comm* newcomm)
{
  *newcomm = g_next_comm++;
  return MPI_SUCCESS;
}

int MPI_Comm_free(MPI_Comm* comm)
{
  *comm = MPI_COMM_NULL;
  return MPI_SUCCESS;
}

int MPI_Comm_split(MPI_Comm, int, int, MPI_Comm* newcomm)
{
  *newcomm = g_next_comm++;
  return MPI_SUCCESS;
}

int MPI_Comm_split_type(MPI_Comm, int, int, MPI_Info, MPI_Comm* newcomm)
{
  *newcomm = g_next_comm++;
  return MPI_SUCCESS;
}

int MPI_Barrier(MPI_Comm)
{
  return MPI_SUCCESS;
}

int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm)
{
  return MPI_SUCCESS;
}

int MPI_Allreduce(const void* sbuf, void* rbuf, int count, MPI_Datatype type, MPI_Op, MPI_Comm)
{
  if (sbuf != MPI_IN_PLACE)
    std::memcpy(rbuf, sbuf, (size_t)count * type);
  return MPI_SUCCESS;
}

int MPI_Reduce(const void* sbuf, void* rbuf, int count, MPI_Datatype type, MPI_Op, int, MPI_Comm)
{
  if (sbuf != MPI_IN_PLACE)
    std::memcpy(rbuf, sbuf, (size_t)count * type);
  return MPI_SUCCESS;
}

int MPI_Allgatherv(const void* sbuf, int scount, MPI_Datatype type, void* rbuf, const int*,
                   const int* displs, MPI_Datatype, MPI_Comm)
{
  if (sbuf != MPI_IN_PLACE)
    std::memcpy((uint8_t*)rbuf + (size_t)displs[0] * type, sbuf, (size_t)scount * type);
  return MPI_SUCCESS;
}

int MPI_Isend(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm,
              MPI_Request* req)
{
  *req = MPI_REQUEST_NULL;
  if (dest == MPI_PROC_NULL)
    return MPI_SUCCESS;
  size_t               bytes = (size_t)count * type;
  std::vector<uint8_t> msg(bytes);
  if (bytes)
    std::memcpy(msg.data(), buf, bytes);
  std::lock_guard<std::mutex> lock(g_mutex);
  g_mailbox[Key(comm, tag)].push_back(std::move(msg));
  return MPI_SUCCESS;
}

int MPI_Irecv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm,
              MPI_Request* req)
{
  *req = MPI_REQUEST_NULL;
  if (source == MPI_PROC_NULL)
    return MPI_SUCCESS;
  std::lock_guard<std::mutex> lock(g_mutex);
  int                         slot;
  if (!g_free_slots.empty()) {
    slot = g_free_slots.back();
    g_free_slots.pop_back();
  } else {
    slot = (int)g_pending.size();
    g_pending.push_back(Pending());
  }
  g_pending[slot] = Pending{buf, count * type, comm, tag, true};
  if (!try_complete(slot))
    *req = slot;
  return MPI_SUCCESS;
}

int MPI_Iprobe(int source, int tag, MPI_Comm comm, int* flag, MPI_Status* status)
{
  if (source == MPI_PROC_NULL) {
    *flag = 1;
    if (status) {
      status->MPI_SOURCE = MPI_PROC_NULL;
      status->MPI_TAG    = tag;
      status->shim_bytes = 0;
    }
    return MPI_SUCCESS;
  }
  std::lock_guard<std::mutex> lock(g_mutex);
  auto                        it = g_mailbox.find(Key(comm, tag));
  if (it == g_mailbox.end() || it->second.empty()) {
    *flag = 0;
  } else {
    *flag = 1;
    if (status) {
      status->MPI_SOURCE = 0;
      status->MPI_TAG    = tag;
      status->shim_bytes = (int)it->second.front().size();
    }
  }
  return MPI_SUCCESS;
}

int MPI_Get_count(const MPI_Status* status, MPI_Datatype type, int* count)
{
  *count = status->shim_bytes / type;
  return MPI_SUCCESS;
}

int MPI_Type_size(MPI_Datatype type, int* size)
{
  *size = type;
  return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request* req, MPI_Status*)
{
  wait_one(req);
  return MPI_SUCCESS;
}

int MPI_Waitall(int n, MPI_Request* reqs, MPI_Status*)
{
  for (int i = 0; i < n; i++)
    wait_one(&reqs[i]);
  return MPI_SUCCESS;
}

int MPI_Testall(int n, MPI_Request* reqs, int* flag, MPI_Status*)
{
  int ok = 1;
  for (int i = 0; i < n; i++)
    ok = ok && test_one(&reqs[i]);
  *flag = ok;
  return MPI_SUCCESS;
}

void picnix_mpi_shim_reset(void)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  g_mailbox.clear();
  g_pending.clear();
  g_free_slots.clear();
}

} // extern "C"
