// Native host packer: flattens the reference's list-of-states rollout batch (TrajBatchDisc:
// design_opt/utils/tools.py:4-16; state layout design_opt/envs/derl_envs/derl_env.py:377-396) into the
// flat fp32 / int32 arrays of the device arena, writing straight into caller-provided (pinned) buffers.
// Replaces the per-state Python work of tensorfy (transform2act_agent.py:20-24) and batch_data
// (transform2act_policy.py:183-202).  Pure CPython C-API + buffer protocol (no numpy headers needed).
//
//   scan(states)  -> (T, n_total, e_total, d_obs, H, O, G)
//   fill(states, actions_or_None, obs, act, node_ptr, edge_ptr, esrc, edst, stage, body, hf, ob, go, nthreads)
// where the outputs are writable buffers (numpy arrays / torch pinned tensors via .numpy()).
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#include <numpy/arrayobject.h>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct Src {
  const void* ptr; Py_ssize_t n0, n1; Py_ssize_t s0, s1; char kind; int itemsize;   // kind: 'f' float, 'i' signed int
};

// numpy arrays (every slot the reference's sampler produces) are read through the array C-API: the buffer protocol
// costs a format-string export per call, ~3x the time of this path over the 450 k arrays of a 50 k-transition batch.
// The array objects stay referenced by the caller's batch for the duration of the packing call.
bool get_src_numpy(PyArrayObject* a, Src& s) {
  const int nd = PyArray_NDIM(a);
  if (nd > 2) return false;
  PyArray_Descr* d = PyArray_DESCR(a);
  char k = 0;
  if (d->kind == 'f') k = 'f';
  else if (d->kind == 'i' || d->kind == 'u' || d->kind == 'b') k = 'i';
  if (!k || !PyArray_ISNOTSWAPPED(a)) return false;
  const int isz = (int)PyArray_ITEMSIZE(a);
  if (k == 'f' && isz != 4 && isz != 8) return false;
  s.ptr = PyArray_DATA(a); s.kind = k; s.itemsize = isz;
  if (nd == 0) { s.n0 = 1; s.n1 = 1; s.s0 = s.s1 = isz; }
  else if (nd == 1) { s.n0 = 1; s.n1 = PyArray_DIM(a, 0); s.s0 = 0; s.s1 = PyArray_STRIDE(a, 0); }
  else { s.n0 = PyArray_DIM(a, 0); s.n1 = PyArray_DIM(a, 1); s.s0 = PyArray_STRIDE(a, 0); s.s1 = PyArray_STRIDE(a, 1); }
  return true;
}

bool get_src(PyObject* o, Src& s, std::vector<Py_buffer>& keep) {
  if (PyArray_Check(o) && get_src_numpy((PyArrayObject*)o, s)) return true;
  Py_buffer b;
  if (PyObject_GetBuffer(o, &b, PyBUF_STRIDES | PyBUF_FORMAT) != 0) return false;
  const char* f = b.format ? b.format : "B";
  while (*f == '<' || *f == '=' || *f == '@' || *f == '|') ++f;
  char k = 0;
  if (*f == 'd' || *f == 'f') k = 'f';
  else if (*f == 'l' || *f == 'q' || *f == 'i' || *f == 'b' || *f == 'h') k = 'i';
  else if (*f == 'L' || *f == 'Q' || *f == 'I' || *f == 'B' || *f == 'H' || *f == '?') k = 'i';
  if (!k || b.ndim > 2) { PyBuffer_Release(&b); PyErr_SetString(PyExc_TypeError, "hostpack: unsupported array dtype / rank"); return false; }
  s.ptr = b.buf; s.kind = k; s.itemsize = (int)b.itemsize;
  if (b.ndim == 0) { s.n0 = 1; s.n1 = 1; s.s0 = s.s1 = b.itemsize; }
  else if (b.ndim == 1) { s.n0 = 1; s.n1 = b.shape[0]; s.s0 = 0; s.s1 = b.strides ? b.strides[0] : b.itemsize; }
  else { s.n0 = b.shape[0]; s.n1 = b.shape[1]; s.s0 = b.strides[0]; s.s1 = b.strides[1]; }
  keep.push_back(b);
  return true;
}

inline double load_f(const Src& s, Py_ssize_t i, Py_ssize_t j) {
  const char* p = (const char*)s.ptr + i * s.s0 + j * s.s1;
  if (s.kind == 'f') return s.itemsize == 8 ? *(const double*)p : (double)*(const float*)p;
  switch (s.itemsize) { case 8: return (double)*(const int64_t*)p; case 4: return (double)*(const int32_t*)p;
                        case 2: return (double)*(const int16_t*)p; default: return (double)*(const int8_t*)p; }
}
inline long long load_i(const Src& s, Py_ssize_t i, Py_ssize_t j) {
  const char* p = (const char*)s.ptr + i * s.s0 + j * s.s1;
  if (s.kind == 'f') return (long long)(s.itemsize == 8 ? *(const double*)p : (double)*(const float*)p);
  switch (s.itemsize) { case 8: return *(const int64_t*)p; case 4: return *(const int32_t*)p;
                        case 2: return *(const int16_t*)p; default: return *(const int8_t*)p; }
}

void copy_f32(const Src& s, float* dst, Py_ssize_t ld) {
  if (s.kind == 'f' && s.itemsize == 8 && s.s1 == 8) {
    for (Py_ssize_t i = 0; i < s.n0; ++i) {
      const double* p = (const double*)((const char*)s.ptr + i * s.s0);
      float* d = dst + i * ld;
      for (Py_ssize_t j = 0; j < s.n1; ++j) d[j] = (float)p[j];
    }
  } else if (s.kind == 'f' && s.itemsize == 4 && s.s1 == 4) {
    for (Py_ssize_t i = 0; i < s.n0; ++i) memcpy(dst + i * ld, (const char*)s.ptr + i * s.s0, (size_t)s.n1 * 4);
  } else {
    for (Py_ssize_t i = 0; i < s.n0; ++i)
      for (Py_ssize_t j = 0; j < s.n1; ++j) dst[i * ld + j] = (float)load_f(s, i, j);
  }
}

struct Out { void* ptr; Py_ssize_t len; Py_buffer b; };
bool get_out(PyObject* o, Out& out, int itemsize, std::vector<Py_buffer>& keep) {
  Py_buffer b;
  if (PyObject_GetBuffer(o, &b, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) != 0) return false;
  if (b.itemsize != itemsize) { PyBuffer_Release(&b); PyErr_SetString(PyExc_TypeError, "hostpack: output buffer itemsize"); return false; }
  out.ptr = b.buf; out.len = b.len / itemsize;
  keep.push_back(b);
  return true;
}

struct Item { Src obs, edges, stage, nn, body, hf, ob, go, act; bool has_act; };

// Fast path of collect(): the batch is a list of lists / tuples of numpy arrays (what the reference's sampler produces).
// The walk only READS object headers (list item pointers, array data pointers, shapes, strides) and touches no reference
// count, so it is split over threads while the calling thread keeps the GIL: the 450 k array headers of a 50 k-transition
// batch are scattered over the heap and the single-threaded walk was bound by those cache misses (10 ms -> ~2 ms).
// Returns false (without setting an error) when anything else is met; the caller then takes the general path.
bool collect_parallel(PyObject* states, PyObject* actions, std::vector<Item>& items, int nthreads) {
  if (!PyList_Check(states)) return false;
  const bool has_act = actions && actions != Py_None;
  if (has_act && !PyList_Check(actions)) return false;
  const Py_ssize_t T = PyList_GET_SIZE(states);
  if (has_act && PyList_GET_SIZE(actions) < T) return false;
  if (T < 4096 || nthreads < 2) return false;
  std::vector<char> bad((size_t)nthreads, 0);
  auto work = [&](int w) {
    const Py_ssize_t lo = T * w / nthreads, hi = T * (w + 1) / nthreads;
    for (Py_ssize_t t = lo; t < hi; ++t) {
      PyObject* st = PyList_GET_ITEM(states, t);
      PyObject** slot_objs = nullptr;
      if (PyList_Check(st) && PyList_GET_SIZE(st) == 8) slot_objs = ((PyListObject*)st)->ob_item;
      else if (PyTuple_Check(st) && PyTuple_GET_SIZE(st) == 8) slot_objs = ((PyTupleObject*)st)->ob_item;
      if (!slot_objs) { bad[(size_t)w] = 1; return; }
      Item& it = items[(size_t)t];
      Src* slots[8] = {&it.obs, &it.edges, &it.stage, &it.nn, &it.body, &it.hf, &it.ob, &it.go};
      for (int k = 0; k < 8; ++k) {
        PyObject* o = slot_objs[k];
        if (!PyArray_Check(o) || !get_src_numpy((PyArrayObject*)o, *slots[k])) { bad[(size_t)w] = 1; return; }
      }
      it.has_act = has_act;
      if (has_act) {
        PyObject* a = PyList_GET_ITEM(actions, t);
        if (!PyArray_Check(a) || !get_src_numpy((PyArrayObject*)a, it.act)) { bad[(size_t)w] = 1; return; }
      }
    }
  };
  std::vector<std::thread> th;
  for (int w = 1; w < nthreads; ++w) th.emplace_back(work, w);
  work(0);
  for (auto& x : th) x.join();
  for (char b : bad) if (b) return false;
  return true;
}

bool collect(PyObject* states, PyObject* actions, std::vector<Item>& items, std::vector<Py_buffer>& keep) {
  const Py_ssize_t T = PySequence_Size(states);
  if (T < 0) return false;
  items.resize((size_t)T);
  {
    unsigned hw = std::thread::hardware_concurrency();
    const int nt = hw >= 16 ? 8 : (hw >= 8 ? 4 : (hw >= 4 ? 2 : 1));
    if (collect_parallel(states, actions, items, nt)) return true;
  }
  const bool has_act = actions && actions != Py_None;
  for (Py_ssize_t t = 0; t < T; ++t) {
    PyObject* st = PySequence_GetItem(states, t);
    if (!st) return false;
    if (PySequence_Size(st) != 8) { Py_DECREF(st); PyErr_SetString(PyExc_ValueError, "hostpack: a state must have 8 slots"); return false; }
    Item& it = items[(size_t)t];
    Src* slots[8] = {&it.obs, &it.edges, &it.stage, &it.nn, &it.body, &it.hf, &it.ob, &it.go};
    for (int k = 0; k < 8; ++k) {
      PyObject* o = PySequence_GetItem(st, k);
      const bool ok = o && get_src(o, *slots[k], keep);
      Py_XDECREF(o);
      if (!ok) { Py_DECREF(st); return false; }
    }
    Py_DECREF(st);
    it.has_act = has_act;
    if (has_act) {
      PyObject* a = PySequence_GetItem(actions, t);
      const bool ok = a && get_src(a, it.act, keep);
      Py_XDECREF(a);
      if (!ok) return false;
    }
  }
  return true;
}

void release(std::vector<Py_buffer>& keep) { for (auto& b : keep) PyBuffer_Release(&b); keep.clear(); }

PyObject* py_scan(PyObject*, PyObject* args) {
  PyObject* states;
  if (!PyArg_ParseTuple(args, "O", &states)) return nullptr;
  std::vector<Item> items; std::vector<Py_buffer> keep;
  if (!collect(states, nullptr, items, keep)) { release(keep); return nullptr; }
  long long n = 0, e = 0;
  for (auto& it : items) { n += it.obs.n0; e += it.edges.n1; }
  const long long d_obs = items.empty() ? 0 : (long long)items[0].obs.n1;
  const long long H = items.empty() ? 0 : (long long)items[0].hf.n1, O = items.empty() ? 0 : (long long)items[0].ob.n1,
                  G = items.empty() ? 0 : (long long)items[0].go.n1;
  release(keep);
  return Py_BuildValue("(LLLLLLL)", (long long)items.size(), n, e, d_obs, H, O, G);
}

// shared by fill() and fill_collected(): items are the collected sources, outs the 11 destination buffers
PyObject* fill_core(const std::vector<Item>& items, PyObject* const o[11], int nthreads) {
  std::vector<Py_buffer> keep;
  Out obs, act, np_, ep, es, ed, stg, body, hf, ob, go;
  Out* outs[11] = {&obs, &act, &np_, &ep, &es, &ed, &stg, &body, &hf, &ob, &go};
  for (int k = 0; k < 11; ++k)
    if (!get_out(o[k], *outs[k], 4, keep)) { release(keep); return nullptr; }
  const size_t T = items.size();
  if (T == 0) { release(keep); Py_RETURN_NONE; }
  const Py_ssize_t D = items[0].obs.n1, H = items[0].hf.n1, O = items[0].ob.n1, G = items[0].go.n1;
  int32_t* node_ptr = (int32_t*)np_.ptr; int32_t* edge_ptr = (int32_t*)ep.ptr;
  long long n = 0, e = 0;
  const char* err = nullptr;
  if (np_.len < (Py_ssize_t)T + 1 || ep.len < (Py_ssize_t)T + 1) { release(keep); PyErr_SetString(PyExc_ValueError, "hostpack: pointer buffers too small"); return nullptr; }
  for (size_t t = 0; t < T; ++t) {
    node_ptr[t] = (int32_t)n; edge_ptr[t] = (int32_t)e;
    const Item& it = items[t];
    if (it.obs.n1 != D || it.hf.n1 != H || it.ob.n1 != O || it.go.n1 != G || it.edges.n0 != 2 || it.body.n1 != it.obs.n0 ||
        (it.has_act && (it.act.n0 != it.obs.n0 || it.act.n1 != 8)))
      err = "hostpack: inconsistent state shapes";
    n += it.obs.n0; e += it.edges.n1;
  }
  node_ptr[T] = (int32_t)n; edge_ptr[T] = (int32_t)e;
  if (!err && (obs.len < n * D || act.len < n * 8 || es.len < e || ed.len < e || body.len < n || (Py_ssize_t)T > stg.len ||
               hf.len < (Py_ssize_t)T * H || ob.len < (Py_ssize_t)T * O || go.len < (Py_ssize_t)T * G))
    err = "hostpack: output buffers too small";
  if (err) { release(keep); PyErr_SetString(PyExc_ValueError, err); return nullptr; }

  auto work = [&](size_t lo, size_t hi) {
    for (size_t t = lo; t < hi; ++t) {
      const Item& it = items[t];
      const long long n0 = node_ptr[t], e0 = edge_ptr[t];
      copy_f32(it.obs, (float*)obs.ptr + n0 * D, D);
      if (it.has_act) copy_f32(it.act, (float*)act.ptr + n0 * 8, 8);
      else memset((float*)act.ptr + n0 * 8, 0, (size_t)it.obs.n0 * 8 * 4);
      for (Py_ssize_t j = 0; j < it.edges.n1; ++j) {
        ((int32_t*)es.ptr)[e0 + j] = (int32_t)load_i(it.edges, 0, j);
        ((int32_t*)ed.ptr)[e0 + j] = (int32_t)load_i(it.edges, 1, j);
      }
      for (Py_ssize_t j = 0; j < it.body.n1; ++j) ((int32_t*)body.ptr)[n0 + j] = (int32_t)load_i(it.body, 0, j);
      ((int32_t*)stg.ptr)[t] = (int32_t)load_i(it.stage, 0, 0);
      copy_f32(it.hf, (float*)hf.ptr + (long long)t * H, H);
      copy_f32(it.ob, (float*)ob.ptr + (long long)t * O, O);
      copy_f32(it.go, (float*)go.ptr + (long long)t * G, G);
    }
  };
  Py_BEGIN_ALLOW_THREADS
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > T) nthreads = (int)T;
  std::vector<std::thread> th;
  const size_t per = (T + nthreads - 1) / nthreads;
  for (int k = 1; k < nthreads; ++k) {
    const size_t lo = k * per, hi = lo + per < T ? lo + per : T;
    if (lo < hi) th.emplace_back(work, lo, hi);
  }
  work(0, per < T ? per : T);
  for (auto& x : th) x.join();
  Py_END_ALLOW_THREADS
  release(keep);
  Py_RETURN_NONE;
}

PyObject* py_fill(PyObject*, PyObject* args) {
  PyObject *states, *actions, *o[11];
  int nthreads = 4;
  if (!PyArg_ParseTuple(args, "OOOOOOOOOOOOO|i", &states, &actions, &o[0], &o[1], &o[2], &o[3], &o[4], &o[5], &o[6], &o[7],
                        &o[8], &o[9], &o[10], &nthreads)) return nullptr;
  std::vector<Item> items; std::vector<Py_buffer> keep;
  if (!collect(states, actions, items, keep)) { release(keep); return nullptr; }
  PyObject* r = fill_core(items, o, nthreads);
  release(keep);
  return r;
}

// One walk over the Python objects instead of two: collect() keeps the acquired buffers in a capsule that
// fill_collected() consumes (the list walk is the serial, GIL-bound part of the host path).
struct Collected { std::vector<Item> items; std::vector<Py_buffer> keep; PyObject* states = nullptr; PyObject* actions = nullptr; };
void collected_free(PyObject* cap) {
  Collected* c = (Collected*)PyCapsule_GetPointer(cap, "sard_b200.collected");
  if (c) { release(c->keep); Py_XDECREF(c->states); Py_XDECREF(c->actions); delete c; }
}
PyObject* py_collect(PyObject*, PyObject* args) {
  PyObject *states, *actions = Py_None;
  if (!PyArg_ParseTuple(args, "O|O", &states, &actions)) return nullptr;
  Collected* c = new Collected();
  if (!collect(states, actions, c->items, c->keep)) { release(c->keep); delete c; return nullptr; }
  // numpy sources are held by raw pointer: keep the batch (and so its arrays) alive as long as the handle
  c->states = states; Py_INCREF(states);
  if (actions != Py_None) { c->actions = actions; Py_INCREF(actions); }
  long long n = 0, e = 0;
  for (auto& it : c->items) { n += it.obs.n0; e += it.edges.n1; }
  const bool any = !c->items.empty();
  const long long T = (long long)c->items.size(), d_obs = any ? (long long)c->items[0].obs.n1 : 0,
                  H = any ? (long long)c->items[0].hf.n1 : 0, O = any ? (long long)c->items[0].ob.n1 : 0,
                  G = any ? (long long)c->items[0].go.n1 : 0;
  PyObject* cap = PyCapsule_New(c, "sard_b200.collected", collected_free);
  if (!cap) { release(c->keep); Py_XDECREF(c->states); Py_XDECREF(c->actions); delete c; return nullptr; }
  return Py_BuildValue("(NLLLLLLL)", cap, T, n, e, d_obs, H, O, G);
}
PyObject* py_fill_collected(PyObject*, PyObject* args) {
  PyObject *cap, *o[11];
  int nthreads = 4;
  if (!PyArg_ParseTuple(args, "OOOOOOOOOOOO|i", &cap, &o[0], &o[1], &o[2], &o[3], &o[4], &o[5], &o[6], &o[7], &o[8], &o[9],
                        &o[10], &nthreads)) return nullptr;
  Collected* c = (Collected*)PyCapsule_GetPointer(cap, "sard_b200.collected");
  if (!c) return nullptr;
  return fill_core(c->items, o, nthreads);
}

PyMethodDef methods[] = {
    {"scan", py_scan, METH_VARARGS, "scan(states) -> (T, n_total, e_total, d_obs, H, O, G)"},
    {"collect", py_collect, METH_VARARGS, "collect(states[, actions]) -> (handle, T, n_total, e_total, d_obs, H, O, G)"},
    {"fill_collected", py_fill_collected, METH_VARARGS, "fill_collected(handle, obs, act, node_ptr, edge_ptr, esrc, edst, stage, body, hf, ob, go[, nthreads])"},
    {"fill", py_fill, METH_VARARGS, "fill(states, actions, obs, act, node_ptr, edge_ptr, esrc, edst, stage, body, hf, ob, go[, nthreads])"},
    {nullptr, nullptr, 0, nullptr}};
PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_hostpack", "native rollout packer for sard_b200", -1, methods};

}  // namespace

PyMODINIT_FUNC PyInit__hostpack(void) {
  import_array();
  return PyModule_Create(&moddef);
}
