// tests/emu/nccl_emu.cpp — TEST INFRASTRUCTURE: an in-process stand-in for the ten NCCL entry points
// capi.cu binds with dlsym, for multi-rank tests on the CPU thread emulation (tests/emu).  Ranks are
// host threads of ONE process (the way qmc4_b200/host/main.cpp --gpus N drives its GPUs); buffers
// are host memory, "streams" are synchronous.  Point QMC4_NCCL_LIB at the built libnccl_emu.so.
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <vector>

namespace {

struct World {
  int nranks = 0, joined = 0;
  std::mutex m;
  std::condition_variable cv;
  // all-reduce rendezvous
  int arrived = 0;
  unsigned phase = 0;
  std::vector<const void *> src;
  // point-to-point mailbox
  struct Msg { int from, to; const void *ptr; size_t bytes; bool done; };
  std::deque<Msg *> box;
};

struct Comm { World *w; int rank; };

std::mutex g_reg_mutex;
std::map<uint64_t, World *> g_worlds;
uint64_t g_next_id = 1;

struct Pending { bool send; void *ptr; size_t bytes; int peer; Comm *c; };
thread_local int t_group_depth = 0;
thread_local std::vector<Pending> t_pending;

size_t elem_size(int dtype) { return dtype == 2 ? 4 : 8; /* ncclInt32 = 2, ncclFloat64 = 8 */ }

void barrier(World *w, std::unique_lock<std::mutex> &lk)
{
  const unsigned ph = w->phase;
  if (++w->arrived == w->nranks)
    {
      w->arrived = 0;
      ++w->phase;
      w->cv.notify_all();
    }
  else w->cv.wait(lk, [&] { return w->phase != ph; });
}

void run_p2p(std::vector<Pending> &ops)
{
  std::vector<World::Msg *> mine;
  for (auto &p : ops)
    if (p.send)
      {
        World *w = p.c->w;
        auto *m = new World::Msg{p.c->rank, p.peer, p.ptr, p.bytes, false};
        std::lock_guard<std::mutex> g(w->m);
        w->box.push_back(m);
        mine.push_back(m);
        w->cv.notify_all();
      }
  for (auto &p : ops)
    if (!p.send)
      {
        World *w = p.c->w;
        std::unique_lock<std::mutex> lk(w->m);
        World::Msg *hit = nullptr;
        w->cv.wait(lk, [&] {
          for (auto *m : w->box)
            if (!m->done && m->from == p.peer && m->to == p.c->rank)
              {
                hit = m;
                return true;
              }
          return false;
        });
        std::memcpy(p.ptr, hit->ptr, hit->bytes < p.bytes ? hit->bytes : p.bytes);
        hit->done = true;
        w->cv.notify_all();
      }
  for (size_t i = 0, k = 0; i < ops.size(); ++i)
    if (ops[i].send)
      {
        World *w = ops[i].c->w;
        World::Msg *m = mine[k++];
        std::unique_lock<std::mutex> lk(w->m);
        w->cv.wait(lk, [&] { return m->done; });
        for (auto it = w->box.begin(); it != w->box.end(); ++it)
          if (*it == m)
            {
              w->box.erase(it);
              break;
            }
        delete m;
      }
  ops.clear();
}

}  // namespace

extern "C" {

typedef struct { char internal[128]; } ncclUniqueId;

int ncclGetUniqueId(ncclUniqueId *id)
{
  std::lock_guard<std::mutex> g(g_reg_mutex);
  std::memset(id, 0, sizeof *id);
  uint64_t v = g_next_id++;
  std::memcpy(id->internal, &v, 8);
  return 0;
}

int ncclCommInitRank(void **comm, int nranks, ncclUniqueId id, int rank)
{
  uint64_t key;
  std::memcpy(&key, id.internal, 8);
  World *w;
  {
    std::lock_guard<std::mutex> g(g_reg_mutex);
    auto it = g_worlds.find(key);
    if (it == g_worlds.end())
      {
        w = new World;
        w->nranks = nranks;
        w->src.resize(nranks);
        g_worlds[key] = w;
      }
    else w = it->second;
  }
  if (w->nranks != nranks || rank < 0 || rank >= nranks) return 4;
  *comm = new Comm{w, rank};
  std::unique_lock<std::mutex> lk(w->m);
  barrier(w, lk);
  return 0;
}

int ncclCommDestroy(void *comm)
{
  delete static_cast<Comm *>(comm);
  return 0;
}

int ncclAllReduce(const void *send, void *recv, size_t count, int dtype, int /*op: sum*/, void *comm, void *)
{
  Comm *c = static_cast<Comm *>(comm);
  World *w = c->w;
  std::vector<unsigned char> tmp(count * elem_size(dtype));
  std::unique_lock<std::mutex> lk(w->m);
  w->src[c->rank] = send;
  barrier(w, lk);
  if (dtype == 8)
    {
      double *t = reinterpret_cast<double *>(tmp.data());
      for (size_t i = 0; i < count; ++i)
        {
          double s = 0;
          for (int r = 0; r < w->nranks; ++r) s += static_cast<const double *>(w->src[r])[i];
          t[i] = s;
        }
    }
  else
    {
      int32_t *t = reinterpret_cast<int32_t *>(tmp.data());
      for (size_t i = 0; i < count; ++i)
        {
          int32_t s = 0;
          for (int r = 0; r < w->nranks; ++r) s += static_cast<const int32_t *>(w->src[r])[i];
          t[i] = s;
        }
    }
  barrier(w, lk);   /* everybody has read every source: in-place results may now be written */
  std::memcpy(recv, tmp.data(), tmp.size());
  barrier(w, lk);
  return 0;
}

int ncclAllGather(const void *send, void *recv, size_t count, int dtype, void *comm, void *)
{
  Comm *c = static_cast<Comm *>(comm);
  World *w = c->w;
  const size_t b = count * elem_size(dtype);
  std::vector<unsigned char> tmp(b * w->nranks);
  std::unique_lock<std::mutex> lk(w->m);
  w->src[c->rank] = send;
  barrier(w, lk);
  for (int r = 0; r < w->nranks; ++r) std::memcpy(tmp.data() + r * b, w->src[r], b);
  barrier(w, lk);
  std::memcpy(recv, tmp.data(), tmp.size());
  barrier(w, lk);
  return 0;
}

int ncclGroupStart()
{
  ++t_group_depth;
  return 0;
}

int ncclGroupEnd()
{
  if (--t_group_depth == 0) run_p2p(t_pending);
  return 0;
}

int ncclSend(const void *buf, size_t count, int dtype, int peer, void *comm, void *)
{
  t_pending.push_back(Pending{true, const_cast<void *>(buf), count * elem_size(dtype), peer, static_cast<Comm *>(comm)});
  if (t_group_depth == 0) run_p2p(t_pending);
  return 0;
}

int ncclRecv(void *buf, size_t count, int dtype, int peer, void *comm, void *)
{
  t_pending.push_back(Pending{false, buf, count * elem_size(dtype), peer, static_cast<Comm *>(comm)});
  if (t_group_depth == 0) run_p2p(t_pending);
  return 0;
}

int ncclCommAbort(void *) { return 0; }

const char *ncclGetErrorString(int e) { return e == 0 ? "no error" : "nccl_emu: invalid argument"; }

}  // extern "C"
