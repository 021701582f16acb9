// host_pool.hpp - a small persistent pool of host threads for the per-node CPU work of a frontier batch (node presolve, marshalling,
// packing of the node images): every node is independent, the device sees the result only once all of them are packed.
// SDPCUDA_HOST_THREADS sets the number of threads (default: the hardware threads divided by the ranks of the node, at most 16);
// 1 = everything on the calling thread.  The calling thread takes part in the work.
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <unistd.h>

namespace sdphost {

class Pool
{
public:
   static constexpr int MAX_WORKERS = 64;            // callers may keep per-worker scratch in arrays of this size
private:
   std::vector<std::thread> workers;
   std::mutex mu, runmu;
   std::condition_variable wake, finished;
   const std::function<void(int, int)>* job = nullptr;
   std::atomic<int> next{0};
   int total = 0, generation = 0, busy = 0;
   bool stopping = false;
   pid_t owner = 0;

   void loop(int worker)
   {
      int seen = 0;
      for( ;; )
      {
         const std::function<void(int, int)>* f;
         {
            std::unique_lock<std::mutex> lk(mu);
            wake.wait(lk, [&] { return stopping || generation != seen; });
            if( stopping ) return;
            seen = generation;
            f = job;
         }
         drain(*f, worker);
         {
            std::lock_guard<std::mutex> lk(mu);
            if( --busy == 0 ) finished.notify_all();
         }
      }
   }
   void drain(const std::function<void(int, int)>& f, int worker)
   {
      for( int i = next.fetch_add(1); i < total; i = next.fetch_add(1) ) f(i, worker);
   }

public:
   static int wanted()
   {
      const char* e = getenv("SDPCUDA_HOST_THREADS");
      if( e != nullptr && atoi(e) > 0 ) return std::min(atoi(e), MAX_WORKERS);
      int hw = (int)std::thread::hardware_concurrency();
      const char* lw = getenv("LOCAL_WORLD_SIZE");
      if( lw != nullptr && atoi(lw) > 1 ) hw /= atoi(lw);
      return std::max(1, std::min(hw, 16));
   }
   static Pool& get() { static Pool* p = new Pool(); return *p; }      // never destroyed: no join at process exit

   void run(int n, const std::function<void(int)>& f) { run_indexed(n, [&](int i, int) { f(i); }); }

   // f(i, worker) for i = 0 ... n-1, each exactly once, on the pool's threads and the caller (worker 0); worker < wanted().
   // Returns when all are done.
   void run_indexed(int n, const std::function<void(int, int)>& f)
   {
      const int nt = wanted();
      if( n < 2 || nt < 2 ) { for( int i = 0; i < n; ++i ) f(i, 0); return; }
      std::lock_guard<std::mutex> serial(runmu);                          // one batch at a time (solver threads of a concurrent SCIP run)
      if( owner != getpid() )                                             // first use, or a forked child (threads do not survive fork)
      {
         if( owner != 0 ) { for( int i = 0; i < n; ++i ) f(i, 0); return; }
         owner = getpid();
         for( int t = 0; t < nt - 1; ++t ) workers.emplace_back([this, t] { loop(t + 1); });
         for( std::thread& t : workers ) t.detach();
      }
      {
         std::lock_guard<std::mutex> lk(mu);
         job = &f; total = n; next.store(0); busy = (int)workers.size(); ++generation;
      }
      wake.notify_all();
      drain(f, 0);
      std::unique_lock<std::mutex> lk(mu);
      finished.wait(lk, [&] { return busy == 0; });
      job = nullptr;
   }
};

} // namespace sdphost
