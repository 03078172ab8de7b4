// multi.cuh -- one frame on several GPUs, behind the C ABI (included by render.cu).
//
// The reference's only way to use more than one processor is to render disjoint `cropwindow`s in separate
// processes and paste the results by data-window offset (film/image.cpp:79-85,148-156,
// tools/exrassemble.cpp:53-66).  Here the film window is cut into pixel tiles, every rank (one GPU each: a
// process of a torchrun job, or a thread of one process) renders its share in tile mode
// (pbrt_b200_render_tiles: a tile's pixels receive exactly the samples they receive in the whole frame, and the
// keyed RNG makes each sample's value independent of who traces it), and ONE exchange moves the finished
// tiles to rank 0: every other rank's tiles leave as a single ncclSend of its contiguous tile block, rank 0
// posts the matching ncclRecv's in one group and pastes all tiles into the frame with one kernel.  No
// arithmetic happens in the exchange, so the assembled frame is bit-identical to the one-GPU frame.
//
// Partition: square tiles of about (pixels / (28 * ranks))^(1/2) on a side, tile (col, row) to rank
// (col + row * shift) % ranks (tile_shift below).  The deal is static on purpose: a wave of a path-traced frame costs ~40 kernel launches
// whose tails run almost empty, so a rank that fetched its work in many small dynamic batches would pay that
// fixed cost per batch (measured: ~0.3 ms per wave against a 20 ms frame share at 8 GPUs); ~30 small tiles per
// rank spread over the whole image balance to within a few per cent and render as ONE wave.  Tile mode culls
// the filter-apron samples that do not reach the tile, so small tiles cost no extra rays.
//
// NCCL is loaded at run time (dlopen "libnccl.so.2"): the library has no link-time dependency on it and
// single-GPU use works where NCCL is absent.
#ifndef PB_MULTI_CUH
#define PB_MULTI_CUH
#include <dlfcn.h>
#include <nccl.h>
#include <atomic>
#include <thread>

namespace {
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};
NcclApi *nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) { api.error = std::string("cannot load NCCL: ") + dlerror(); return; }
#define PB_NCCL_SYM(field, sym) \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, sym)); \
  if (!api.field) api.error = std::string("NCCL symbol missing: ") + sym;
    PB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    PB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    PB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    PB_NCCL_SYM(Send, "ncclSend")
    PB_NCCL_SYM(Recv, "ncclRecv")
    PB_NCCL_SYM(GroupStart, "ncclGroupStart")
    PB_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    PB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef PB_NCCL_SYM
  });
  return &api;
}
#define NCCL_TRY(expr)                                                                                  \
  do {                                                                                                  \
    ncclResult_t r_ = (expr);                                                                           \
    if (r_ != ncclSuccess)                                                                              \
      return fail(PBRT_B200_ERR_CUDA, std::string(#expr) + ": " + nccl_api()->GetErrorString(r_));       \
  } while (0)

struct TileRef {          // one tile of the frame and where its resolved pixels sit
  int x0, y0, w, h;
  const float4 *src;
};
// rank 0: every tile of every rank into the frame (row-major float4, frame window origin fx0, fy0)
__global__ void k_paste_tiles(const TileRef *tiles, int nTiles, float4 *frame, int fx0, int fy0, int frameW) {
  for (int t = blockIdx.y; t < nTiles; t += gridDim.y) {
    const TileRef tr = tiles[t];
    const int n = tr.w * tr.h;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      int y = i / tr.w, x = i - y * tr.w;
      frame[(size_t)(tr.y0 - fy0 + y) * frameW + (tr.x0 - fx0 + x)] = tr.src[i];
    }
  }
}
}  // namespace

struct PbrtB200Comm {
  int rank, world, device;
  ncclComm_t comm;
};

// The partition, a pure function of the frame window and the number of ranks.
static int frame_tile_edge(const PbrtB200Window &f, int world, int requested) {
  if (requested > 0) return requested;
  double target = sqrt((double)f.width * f.height / (28.0 * std::max(1, world)));
  int edge = (int)(target / 16.0 + 0.5) * 16;
  return std::max(32, std::min(256, edge));
}
static void frame_tiles(const PbrtB200Window &f, int edge, std::vector<PbrtB200Window> *tiles, int *cols) {
  tiles->clear();
  *cols = (f.width + edge - 1) / edge;
  for (int ty = f.y0; ty < f.y0 + f.height; ty += edge)
    for (int tx = f.x0; tx < f.x0 + f.width; tx += edge) {
      PbrtB200Window t = {tx, ty, std::min(edge, f.x0 + f.width - tx), std::min(edge, f.y0 + f.height - ty)};
      tiles->push_back(t);
    }
}
// Owner of tile (col, row): (col + row * shift) mod world, a lattice.  Every rank's tiles are spread evenly over the
// image in both directions when shift is coprime to world and near world * 0.618 (for 8 ranks: 5).  The round-1 rule
// (i + i / cols) mod world is this lattice with shift = cols + 1, which degenerates to vertical stripes whenever
// world divides cols + 1: 4 ranks on a 1024-pixel frame (11 columns of 96-pixel tiles) left one rank with 2 of the
// 11 columns -- 23.9 ms of work beside 39.3 ms on its neighbour (profiles/r02s_scale_n4.json).
static inline int tile_shift(int world) {
  if (world <= 2) return 1;
  auto gcd = [](int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; };
  int s = (int)(world * 0.6180339887 + 0.5);
  for (int d = 0; d < world; ++d) {
    if (s + d < world && gcd(s + d, world) == 1) return s + d;
    if (s - d > 0 && gcd(s - d, world) == 1) return s - d;
  }
  return 1;
}
static inline int tile_owner(int i, int cols, int world) { return (i % cols + (i / cols) * tile_shift(world)) % world; }

extern "C" {

// The partition pbrt_b200_render_frame uses, for callers and tests (host arithmetic only, no device needed):
// tiles in raster order and the rank each belongs to.  Returns the tile count (also when it exceeds max_tiles;
// only the first max_tiles entries are written then).
int pbrt_b200_frame_partition(const PbrtB200Window *frame, int world, int tile_edge, PbrtB200Window *tiles_out,
                              int32_t *owner_out, int max_tiles) {
  if (!frame || world < 1 || frame->width < 1 || frame->height < 1) return 0;
  std::vector<PbrtB200Window> tiles;
  int cols = 0;
  frame_tiles(*frame, frame_tile_edge(*frame, world, tile_edge), &tiles, &cols);
  for (size_t i = 0; i < tiles.size() && (int)i < max_tiles; ++i) {
    if (tiles_out) tiles_out[i] = tiles[i];
    if (owner_out) owner_out[i] = tile_owner((int)i, cols, world);
  }
  return (int)tiles.size();
}

int pbrt_b200_comm_unique_id(void *id128) {
  NcclApi *api = nccl_api();
  if (!api->error.empty()) return fail(PBRT_B200_ERR_UNSUPPORTED, api->error);
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  NCCL_TRY(api->GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return PBRT_B200_OK;
}

int pbrt_b200_comm_create(const void *id128, int rank, int world, int device, PbrtB200Comm **out) {
  if (!id128 || !out || world < 1 || rank < 0 || rank >= world) return fail(PBRT_B200_ERR_INVALID, "bad communicator arguments");
  *out = nullptr;
  NcclApi *api = nccl_api();
  if (!api->error.empty()) return fail(PBRT_B200_ERR_UNSUPPORTED, api->error);
  CU_TRY(cudaSetDevice(device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  PbrtB200Comm *c = new PbrtB200Comm;
  c->rank = rank; c->world = world; c->device = device; c->comm = nullptr;
  ncclResult_t r = api->CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    delete c;
    return fail(PBRT_B200_ERR_CUDA, std::string("ncclCommInitRank: ") + api->GetErrorString(r));
  }
  *out = c;
  return PBRT_B200_OK;
}

void pbrt_b200_comm_destroy(PbrtB200Comm *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->comm) nccl_api()->CommDestroy(c->comm);
  delete c;
}

int pbrt_b200_render_frame(PbrtB200Handle *h, PbrtB200Comm *comm, int tile_edge, float *rgba_host,
                           PbrtB200Stats *stats) {
  if (!h) return fail(PBRT_B200_ERR_INVALID, "null handle");
  const int rank = comm ? comm->rank : 0, world = comm ? comm->world : 1;
  if (comm && comm->device != h->device) return fail(PBRT_B200_ERR_INVALID, "communicator and scene live on different devices");
  CU_TRY(cudaSetDevice(h->device));
  const PbrtB200Window &f = h->frameWin;
  PbrtB200Stats st;
  memset(&st, 0, sizeof(st));
  if (world == 1) {
    // the whole frame as one window (crop semantics == tile semantics for the film window itself)
    int rc = render_windows(h, &f, 1, &st, false);
    if (rc != PBRT_B200_OK) return rc;
    if (rgba_host) CU_TRY(cudaMemcpy(rgba_host, h->films[0].rgba, (size_t)f.width * f.height * sizeof(float4), cudaMemcpyDeviceToHost));
    h->frame = h->films[0].rgba;
    if (stats) *stats = st;
    return PBRT_B200_OK;
  }
  NcclApi *api = nccl_api();
  std::vector<PbrtB200Window> tiles;
  int cols = 0;
  const int edge = frame_tile_edge(f, world, tile_edge);
  frame_tiles(f, edge, &tiles, &cols);
  std::vector<PbrtB200Window> mine;
  std::vector<size_t> pixelsOf(world, 0);
  for (size_t i = 0; i < tiles.size(); ++i) {
    int o = tile_owner((int)i, cols, world);
    pixelsOf[o] += (size_t)tiles[i].width * tiles[i].height;
    if (o == rank) mine.push_back(tiles[i]);
  }
  if (!mine.empty()) {
    int rc = render_windows(h, mine.data(), (int)mine.size(), &st, true);
    if (rc != PBRT_B200_OK) return rc;
  }
  cudaEvent_t e0, e1;
  CU_TRY(cudaEventCreate(&e0)); CU_TRY(cudaEventCreate(&e1));
  CU_TRY(cudaEventRecord(e0, h->stream));
  const size_t framePixels = (size_t)f.width * f.height;
  if (rank != 0) {
    if (!mine.empty()) NCCL_TRY(api->Send(h->rgbaSlab, pixelsOf[rank] * 4, ncclFloat, 0, comm->comm, h->stream));
  } else {
    // staging for the other ranks' tile blocks + the frame, kept with the handle between frames
    size_t remote = framePixels - pixelsOf[0];
    if (h->gatherFloats < (remote + framePixels) * 4) {
      pool_free(h->device, h->gather);
      h->gather = nullptr; h->gatherFloats = 0;
      CU_TRY(pool_alloc(h->device, (void **)&h->gather, (remote + framePixels) * sizeof(float4)));
      h->gatherFloats = (remote + framePixels) * 4;
    }
    float4 *frame = h->gather, *staging = h->gather + framePixels;
    std::vector<float4 *> base(world, nullptr);
    base[0] = h->rgbaSlab;
    size_t off = 0;
    NCCL_TRY(api->GroupStart());
    for (int r = 1; r < world; ++r) {
      base[r] = staging + off;
      if (pixelsOf[r]) NCCL_TRY(api->Recv(base[r], pixelsOf[r] * 4, ncclFloat, r, comm->comm, h->stream));
      off += pixelsOf[r];
    }
    NCCL_TRY(api->GroupEnd());
    // tile table: every tile's source is its position inside its owner's block (blocks hold tiles in raster order)
    std::vector<TileRef> table(tiles.size());
    std::vector<size_t> used(world, 0);
    for (size_t i = 0; i < tiles.size(); ++i) {
      int o = tile_owner((int)i, cols, world);
      TileRef tr = {tiles[i].x0, tiles[i].y0, tiles[i].width, tiles[i].height, base[o] + used[o]};
      used[o] += (size_t)tiles[i].width * tiles[i].height;
      table[i] = tr;
    }
    if (h->tileTableCap < table.size()) {
      pool_free(h->device, h->tileTable);
      h->tileTable = nullptr; h->tileTableCap = 0;
      CU_TRY(pool_alloc(h->device, &h->tileTable, table.size() * sizeof(TileRef)));
      h->tileTableCap = table.size();
    }
    CU_TRY(cudaMemcpyAsync(h->tileTable, table.data(), table.size() * sizeof(TileRef), cudaMemcpyHostToDevice, h->stream));
    dim3 grid((unsigned)std::max(1, (edge * edge + 255) / 256), (unsigned)std::min<size_t>(table.size(), 65535));
    k_paste_tiles<<<grid, 256, 0, h->stream>>>((const TileRef *)h->tileTable, (int)table.size(), frame, f.x0, f.y0, f.width);
    h->frame = frame;
    ++st.kernel_launches;
  }
  CU_TRY(cudaEventRecord(e1, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));   // (the host-side tile table must outlive its copy)
  CU_TRY(cudaGetLastError());
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  st.download_ms = ms;   // the exchange + paste, on the device
  if (rank == 0 && rgba_host) {
    CU_TRY(cudaMemcpy(rgba_host, h->frame, framePixels * sizeof(float4), cudaMemcpyDeviceToHost));
    st.d2h_bytes = framePixels * sizeof(float4);
  }
  if (stats) *stats = st;
  return PBRT_B200_OK;
}

int pbrt_b200_frame_device_ptr(PbrtB200Handle *h, void **dptr, uint64_t *n_floats) {
  if (!h || !h->frame) return fail(PBRT_B200_ERR_INVALID, "no assembled frame on this rank");
  *dptr = h->frame;
  *n_floats = (uint64_t)h->frameWin.width * h->frameWin.height * 4;
  return PBRT_B200_OK;
}

// The host-buffer call for several GPUs of this process (what Scene::Render makes when PBRT_GPUS > 1): one
// thread per device, each uploads the scene, renders its tiles and takes part in the exchange; the frame
// arrives in `rgba_out` (film window, row-major RGBA).  Statistics are summed over the devices.
int pbrt_b200_render_host_multi(const PbrtB200Scene *scene, const int *devices, int n_devices, float *rgba_out,
                                PbrtB200Stats *stats) {
  if (!scene || !devices || n_devices < 1) return fail(PBRT_B200_ERR_INVALID, "bad arguments");
  if (n_devices == 1 || scene->sampler.kind == PBRT_B200_SAMPLER_BESTCANDIDATE) {
    // (bestcandidate: its table placements start at the window, so tiles would not equal the frame; one device)
    PbrtB200Window w = {scene->film.x_pixel_start, scene->film.y_pixel_start, scene->film.x_pixel_count, scene->film.y_pixel_count};
    float *outs[1] = {rgba_out};
    return pbrt_b200_render_host(scene, devices[0], &w, 1, outs, nullptr, stats);
  }
  uint8_t id[128];
  int rc = pbrt_b200_comm_unique_id(id);
  if (rc != PBRT_B200_OK) return rc;
  std::vector<int> codes(n_devices, PBRT_B200_OK);
  std::vector<std::string> errors(n_devices);
  std::vector<PbrtB200Stats> sts(n_devices);
  std::vector<std::thread> threads;
  for (int r = 0; r < n_devices; ++r)
    threads.emplace_back([&, r] {
      PbrtB200Handle *h = nullptr;
      PbrtB200Comm *c = nullptr;
      memset(&sts[r], 0, sizeof(PbrtB200Stats));
      // (communicator first: every rank must reach ncclCommInitRank even if its scene upload fails later)
      int code = pbrt_b200_comm_create(id, r, n_devices, devices[r], &c);
      if (code == PBRT_B200_OK) code = pbrt_b200_scene_create(scene, devices[r], &h);
      if (code == PBRT_B200_OK) {
        sts[r].h2d_bytes = h->sceneBytes;
        uint64_t h2d = h->sceneBytes;
        code = pbrt_b200_render_frame(h, c, 0, r == 0 ? rgba_out : nullptr, &sts[r]);
        sts[r].h2d_bytes = h2d;
      }
      if (code != PBRT_B200_OK) errors[r] = g_lastError;
      codes[r] = code;
      if (h) pbrt_b200_scene_destroy(h);
      if (c) pbrt_b200_comm_destroy(c);
    });
  for (auto &t : threads) t.join();
  for (int r = 0; r < n_devices; ++r)
    if (codes[r] != PBRT_B200_OK) return fail(codes[r], "device " + std::to_string(devices[r]) + ": " + errors[r]);
  if (stats) {
    PbrtB200Stats total;
    memset(&total, 0, sizeof(total));
    for (int r = 0; r < n_devices; ++r) {
      const PbrtB200Stats &s = sts[r];
      total.camera_rays += s.camera_rays; total.intersect_rays += s.intersect_rays; total.shadow_rays += s.shadow_rays;
      total.points_shaded += s.points_shaded; total.interior_visits += s.interior_visits; total.leaf_visits += s.leaf_visits;
      total.prim_refs_read += s.prim_refs_read; total.prim_tests += s.prim_tests;
      total.nan_samples += s.nan_samples; total.negative_samples += s.negative_samples; total.infinite_samples += s.infinite_samples;
      total.render_ms = std::max(total.render_ms, s.render_ms);
      total.download_ms = std::max(total.download_ms, s.download_ms);
      total.kernel_launches += s.kernel_launches;
      total.h2d_bytes += s.h2d_bytes; total.d2h_bytes += s.d2h_bytes;
    }
    *stats = total;
  }
  return PBRT_B200_OK;
}

}  // extern "C"
#endif
