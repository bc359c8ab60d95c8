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

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

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

// derived datatypes: one contiguous block of `bytes` at byte offset `offset` inside the file view
struct Derived {
  long long bytes;
  long long offset;
};
std::vector<Derived> g_types;

long long type_bytes(MPI_Datatype t)
{
  if (t < PICNIX_SHIM_DERIVED_BASE)
    return t;
  std::lock_guard<std::mutex> lock(g_mutex);
  return g_types[t - PICNIX_SHIM_DERIVED_BASE].bytes;
}

long long type_offset(MPI_Datatype t)
{
  if (t < PICNIX_SHIM_DERIVED_BASE)
    return 0;
  std::lock_guard<std::mutex> lock(g_mutex);
  return g_types[t - PICNIX_SHIM_DERIVED_BASE].offset;
}

MPI_Datatype new_type(long long bytes, long long offset)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  g_types.push_back(Derived{bytes, offset});
  return PICNIX_SHIM_DERIVED_BASE + (int)g_types.size() - 1;
}

struct ShimFile {
  int       fd = -1;
  long long pos = 0;        // individual file pointer, bytes
  long long view_disp = 0;  // MPI_File_set_view: displacement
  long long view_etype = 1; //                    etype size
  long long view_block = 0; //                    offset of this process' block inside the filetype
};
std::vector<ShimFile> g_files;

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

int MPI_Allgather(const void* sbuf, int scount, MPI_Datatype type, void* rbuf, int, MPI_Datatype, MPI_Comm)
{
  if (sbuf != MPI_IN_PLACE)
    std::memcpy(rbuf, sbuf, (size_t)scount * type_bytes(type));
  return MPI_SUCCESS;
}

int MPI_Gather(const void* sbuf, int scount, MPI_Datatype type, void* rbuf, int, MPI_Datatype, int, MPI_Comm)
{
  if (sbuf != MPI_IN_PLACE)
    std::memcpy(rbuf, sbuf, (size_t)scount * type_bytes(type));
  return MPI_SUCCESS;
}

int MPI_Gatherv(const void* sbuf, int scount, MPI_Datatype type, void* rbuf, const int*, const int* displs,
                MPI_Datatype, int, MPI_Comm)
{
  if (sbuf != MPI_IN_PLACE)
    std::memcpy((uint8_t*)rbuf + (size_t)(displs ? displs[0] : 0) * type_bytes(type), sbuf,
                (size_t)scount * type_bytes(type));
  return MPI_SUCCESS;
}

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype* newtype)
{
  *newtype = new_type((long long)count * type_bytes(oldtype), 0);
  return MPI_SUCCESS;
}

int MPI_Type_create_hindexed(int count, const int* blocklens, const MPI_Aint* displs, MPI_Datatype oldtype,
                             MPI_Datatype* newtype)
{
  if (count != 1) {
    std::fprintf(stderr, "[mpi_shim] MPI_Type_create_hindexed: only one block is supported\n");
    std::abort();
  }
  *newtype = new_type((long long)blocklens[0] * type_bytes(oldtype), displs[0]);
  return MPI_SUCCESS;
}

int MPI_Type_create_subarray(int ndim, const int* gshape, const int* lshape, const int* offset, int,
                             MPI_Datatype oldtype, MPI_Datatype* newtype)
{
  long long n = 1;
  for (int i = 0; i < ndim; i++) {
    if (gshape[i] != lshape[i] || offset[i] != 0) {
      std::fprintf(stderr, "[mpi_shim] MPI_Type_create_subarray: one process owns the whole array\n");
      std::abort();
    }
    n *= lshape[i];
  }
  *newtype = new_type(n * type_bytes(oldtype), 0);
  return MPI_SUCCESS;
}

int MPI_Type_commit(MPI_Datatype*)
{
  return MPI_SUCCESS;
}

int MPI_Type_free(MPI_Datatype*)
{
  return MPI_SUCCESS; // handles are never reused: sizes stay valid for requests in flight
}

int MPI_File_open(MPI_Comm, const char* filename, int amode, MPI_Info, MPI_File* fh)
{
  int flags = 0;
  if (amode & MPI_MODE_RDWR)
    flags |= O_RDWR;
  else if (amode & MPI_MODE_WRONLY)
    flags |= O_WRONLY;
  else
    flags |= O_RDONLY;
  if (amode & MPI_MODE_CREATE)
    flags |= O_CREAT;
  const int fd = ::open(filename, flags, 0644);
  if (fd < 0)
    return 1;
  std::lock_guard<std::mutex> lock(g_mutex);
  ShimFile                    f;
  f.fd = fd;
  g_files.push_back(f);
  *fh = (int)g_files.size() - 1;
  return MPI_SUCCESS;
}

int MPI_File_close(MPI_File* fh)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  if (*fh >= 0 && *fh < (int)g_files.size() && g_files[*fh].fd >= 0) {
    ::close(g_files[*fh].fd);
    g_files[*fh].fd = -1;
  }
  *fh = -1;
  return MPI_SUCCESS;
}

int MPI_File_delete(const char* filename, MPI_Info)
{
  return ::unlink(filename) == 0 ? MPI_SUCCESS : 1;
}

int MPI_File_seek(MPI_File fh, MPI_Offset offset, int whence)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  ShimFile&                   f = g_files[fh];
  if (whence == MPI_SEEK_SET)
    f.pos = offset * f.view_etype;
  else if (whence == MPI_SEEK_CUR)
    f.pos += offset * f.view_etype;
  else {
    struct stat st;
    ::fstat(f.fd, &st);
    f.pos = st.st_size - f.view_disp + offset * f.view_etype;
  }
  return MPI_SUCCESS;
}

int MPI_File_get_size(MPI_File fh, MPI_Offset* size)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  struct stat                 st;
  ::fstat(g_files[fh].fd, &st);
  *size = st.st_size;
  return MPI_SUCCESS;
}

int MPI_File_get_position(MPI_File fh, MPI_Offset* pos)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  *pos = g_files[fh].pos / g_files[fh].view_etype;
  return MPI_SUCCESS;
}

int MPI_File_set_view(MPI_File fh, MPI_Offset disp, MPI_Datatype etype, MPI_Datatype filetype, const char*,
                      MPI_Info)
{
  const long long es = type_bytes(etype), bo = type_offset(filetype);
  std::lock_guard<std::mutex> lock(g_mutex);
  ShimFile&                   f = g_files[fh];
  f.view_disp  = disp;
  f.view_etype = es;
  f.view_block = bo;
  f.pos        = 0;
  return MPI_SUCCESS;
}

namespace
{
int file_rw(MPI_File fh, long long where, void* buf, long long bytes, bool write)
{
  int fd;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    fd = g_files[fh].fd;
  }
  uint8_t*  p    = static_cast<uint8_t*>(buf);
  long long done = 0;
  while (done < bytes) {
    const ssize_t n = write ? ::pwrite(fd, p + done, bytes - done, where + done)
                            : ::pread(fd, p + done, bytes - done, where + done);
    if (n <= 0)
      return 1;
    done += n;
  }
  return MPI_SUCCESS;
}
} // namespace

int MPI_File_iread_all(MPI_File fh, void* buf, int count, MPI_Datatype type, MPI_Request* req)
{
  *req = MPI_REQUEST_NULL;
  long long where;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    where = g_files[fh].view_disp + g_files[fh].view_block;
  }
  return file_rw(fh, where, buf, (long long)count * type_bytes(type), false);
}

int MPI_File_iwrite_all(MPI_File fh, const void* buf, int count, MPI_Datatype type, MPI_Request* req)
{
  *req = MPI_REQUEST_NULL;
  long long where;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    where = g_files[fh].view_disp + g_files[fh].view_block;
  }
  return file_rw(fh, where, const_cast<void*>(buf), (long long)count * type_bytes(type), true);
}

int MPI_File_iread_at(MPI_File fh, MPI_Offset offset, void* buf, int count, MPI_Datatype type, MPI_Request* req)
{
  *req = MPI_REQUEST_NULL;
  long long where;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    where = g_files[fh].view_disp + offset * g_files[fh].view_etype;
  }
  return file_rw(fh, where, buf, (long long)count * type_bytes(type), false);
}

int MPI_File_iwrite_at(MPI_File fh, MPI_Offset offset, const void* buf, int count, MPI_Datatype type,
                       MPI_Request* req)
{
  *req = MPI_REQUEST_NULL;
  long long where;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    where = g_files[fh].view_disp + offset * g_files[fh].view_etype;
  }
  return file_rw(fh, where, const_cast<void*>(buf), (long long)count * type_bytes(type), true);
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
