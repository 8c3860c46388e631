"""Python mirror of the reference's calling interface (call.hpp:12-43) over the C ABI.

The four reference functions take a std::istream and return std::vector<OutputRecord>; here they
take the pileup text (bytes / bytearray / numpy uint8) and return a list of OutputRecord, with the
same argument meaning and the same error behaviour (a malformed line raises MalformedPileup, the
std::invalid_argument of pileup.cpp:9-10).  All work happens in libsidgpu.so on the GPU.
"""
import collections
import ctypes

import numpy as np

from . import _lib
from ._lib import Columns, Config, Fit, Params, SitesView, UniqueView

METHODS = {"local": 0, "bayes": 1, "likelihood_ratio": 2, "quality": 3}
CSV_HEADER = b"chrom,pos,label,gt,hom_conf,het_conf,conf_type\n"      # sid.cpp:102

OutputRecord = collections.namedtuple(
    "OutputRecord", "chromosome_name position label genotype confidence_homozygous confidence_heterozygous confidence_type")


class SidGpuError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("sidgpu error %d: %s" % (code, message))
        self.code = code


class MalformedPileup(ValueError):
    """std::invalid_argument{"Malformed pileup line"} (pileup.cpp:9-10,22-62)."""


_MALFORMED_CODES = (3, 4, 5)


class DeviceBuffer:
    def __init__(self, ctx, nbytes):
        self.ctx = ctx
        self.nbytes = int(nbytes)
        p = ctypes.c_void_p()
        ctx._ck(ctx.lib.sidgpu_malloc(ctx.h, self.nbytes, ctypes.byref(p)))
        self.ptr = p.value

    def free(self):
        if self.ptr:
            self.ctx.lib.sidgpu_free(self.ctx.h, self.ptr)
            self.ptr = None

    def upload(self, array):
        a = np.ascontiguousarray(array)
        assert a.nbytes <= self.nbytes
        self.ctx._ck(self.ctx.lib.sidgpu_memcpy_h2d(self.ctx.h, self.ptr, a.ctypes.data, a.nbytes))
        return self

    def download(self, dtype, count):
        out = np.empty(int(count), dtype=dtype)
        if out.nbytes:
            self.ctx._ck(self.ctx.lib.sidgpu_memcpy_d2h(self.ctx.h, out.ctypes.data, self.ptr, out.nbytes))
        return out


def _as_u8(text):
    if isinstance(text, np.ndarray):
        return np.ascontiguousarray(text.view(np.uint8).reshape(-1))
    if isinstance(text, str):
        text = text.encode()
    return np.frombuffer(bytes(text) if not isinstance(text, (bytes, bytearray, memoryview)) else text, dtype=np.uint8)


class Context:
    """One GPU context (sidgpu_ctx).  Not thread safe; one per GPU."""

    def __init__(self, device=0, max_chunk_bytes=0, max_sites=0, table_log2=0, stream=None):
        self.lib = _lib.load()
        cfg = Config(device, max_chunk_bytes, max_sites, table_log2, stream)
        h = ctypes.c_void_p()
        rc = self.lib.sidgpu_create(ctypes.byref(cfg), ctypes.byref(h))
        if rc != 0:
            raise SidGpuError(rc, self.lib.sidgpu_last_error(None).decode())
        self.h = h

    def close(self):
        if self.h:
            self.lib.sidgpu_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            msg = self.lib.sidgpu_last_error(self.h).decode()
            if rc in _MALFORMED_CODES:
                raise MalformedPileup(msg)
            raise SidGpuError(rc, msg)

    # ---- memory helpers
    def device_buffer(self, nbytes):
        return DeviceBuffer(self, nbytes)

    def upload_text(self, text):
        a = _as_u8(text)
        buf = DeviceBuffer(self, ((a.nbytes + 15) // 16 + 1) * 16)
        if a.nbytes:
            buf.upload(a)
        buf.text_len = a.nbytes
        return buf

    def copy_d2d(self, dst_ptr, src_ptr, nbytes, sync=True):
        if nbytes:
            f = self.lib.sidgpu_memcpy_d2d if sync else self.lib.sidgpu_memcpy_d2d_async
            self._ck(f(self.h, dst_ptr, src_ptr, nbytes))

    def _download(self, ptr, dtype, count):
        out = np.empty(int(count), dtype=dtype)
        if out.nbytes:
            self._ck(self.lib.sidgpu_memcpy_d2h(self.h, out.ctypes.data, ptr, out.nbytes))
        return out

    @property
    def launch_count(self):
        return int(self.lib.sidgpu_launch_count(self.h))

    def profile(self, enable=True):
        self._ck(self.lib.sidgpu_profile(self.h, 1 if enable else 0))

    def kernel_times(self):
        """{'tokenize': (ms, launches), 'classify': ..., 'csv': ..., 'order': ..., 'fit': ..., 'histogram': ..., 'quality': ...,
        'inflate': ...}
        since profile(True)."""
        ms = (ctypes.c_double * 8)()
        n = (ctypes.c_uint64 * 8)()
        self._ck(self.lib.sidgpu_kernel_times(self.h, ms, n))
        return {k: (ms[i], n[i]) for i, k in enumerate(("tokenize", "classify", "csv", "order", "fit", "histogram", "quality", "inflate"))}

    # ---- K1
    def tokenize(self, d_text, text_len, begin=0, end=None, want_qual=False, strands=False):
        """parsePileupLine/parseReadBases over device text; results copied back as numpy arrays.
        strands: also the per-strand profiles of every site ("fwd", "rev": packed like "profile", fwd + rev == profile;
        the strands of pileup.hpp:15 summed per site)."""
        end = text_len if end is None else end
        v = SitesView()
        ptr = d_text.ptr if isinstance(d_text, DeviceBuffer) else d_text
        self._ck(self.lib.sidgpu_tokenize(self.h, ptr, text_len, begin, end, 1 if want_qual else (3 if strands else 0), ctypes.byref(v)))
        n = v.n_sites
        fwd = rev = None
        if strands and want_qual:
            # the quality columns are wanted (and validated) too: the strands by their own pass over the line offsets
            buf = DeviceBuffer(self, max(16, 16 * n))
            try:
                self._ck(self.lib.sidgpu_strand_counts(self.h, ptr, text_len, v.d_line_off, n, buf.ptr, buf.ptr + 8 * n))
                both = buf.download(np.uint64, 2 * n)
                fwd, rev = both[:n].copy(), both[n:].copy()
            finally:
                buf.free()
        elif strands:
            # a by-product of the tokenizer pass: the forward profile; the reverse one is the rest of the profile, count by count
            fwd = self._download(v.d_fwd, np.uint64, n)
            prof = self._download(v.d_profile, np.uint64, n)
            f4 = np.stack([(fwd >> np.uint64(16 * i)) & np.uint64(0xFFFF) for i in range(4)], axis=1)
            p4 = np.stack([(prof >> np.uint64(16 * i)) & np.uint64(0xFFFF) for i in range(4)], axis=1)
            r4 = (p4 - f4) & np.uint64(0xFFFF)
            rev = r4[:, 0] | (r4[:, 1] << np.uint64(16)) | (r4[:, 2] << np.uint64(32)) | (r4[:, 3] << np.uint64(48))
        names_pool = self._download(v.d_names, np.uint8, v.names_bytes).tobytes()
        refs = self._download(v.d_name_ref, np.uint32, n)
        cache = {}

        def name_of(r):
            if r not in cache:
                ln = names_pool[r] | (names_pool[r + 1] << 8)
                cache[r] = names_pool[r + 2:r + 2 + ln].decode("latin-1")
            return cache[r]

        return {
            "n_sites": n,
            "profile": self._download(v.d_profile, np.uint64, n),
            "pos": self._download(v.d_pos, np.int32, n),
            "slot": self._download(v.d_slot, np.uint32, n),
            "line_off": self._download(v.d_line_off, np.uint64, n) if (want_qual or strands) else None,
            "fwd": fwd,
            "rev": rev,
            "chrom": [name_of(int(r)) for r in refs],
        }

    # ---- sessions
    @staticmethod
    def make_params(method, estimate_prior=False, prior=-1.0, error_threshold=0.1, significance_level=0.05, fit=None,
                    het_only=False, strands=False):
        p = Params()
        p.method = METHODS[method] if isinstance(method, str) else int(method)
        p.estimate_prior = 1 if estimate_prior else 0
        p.prior = prior
        p.error_threshold = error_threshold
        p.significance_level = significance_level
        p.het_only = 1 if het_only else 0
        p.want_strands = 1 if strands else 0
        if fit is not None:
            p.fit_given = 1
            p.fit_pi, p.fit_eps = fit[0], fit[1]
            for i in range(4):
                p.fit_nd[i] = fit[2][i]
        return p

    def begin(self, params):
        self._ck(self.lib.sidgpu_begin(self.h, ctypes.byref(params)))

    def feed(self, d_text, text_len, begin=0, end=None):
        end = text_len if end is None else end
        n = ctypes.c_uint64()
        ptr = d_text.ptr if isinstance(d_text, DeviceBuffer) else d_text
        self._ck(self.lib.sidgpu_feed(self.h, ptr, text_len, begin, end, ctypes.byref(n)))
        return n.value

    def feed_rows(self, d_text, text_len, d_out, out_cap, begin=0, end=None):
        """Text in, CSV rows out in one kernel pass (`local` without -R): returns (csv_bytes, rows, n_sites)."""
        end = text_len if end is None else end
        b, r, n = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        ptr = d_text.ptr if isinstance(d_text, DeviceBuffer) else d_text
        out = d_out.ptr if isinstance(d_out, DeviceBuffer) else d_out
        self._ck(self.lib.sidgpu_feed_rows(self.h, ptr, text_len, begin, end, out, out_cap, ctypes.byref(b), ctypes.byref(r), ctypes.byref(n)))
        return b.value, r.value, n.value

    def finish(self):
        self._ck(self.lib.sidgpu_finish(self.h))

    def emit_csv(self, site_begin, n_sites, d_out, out_cap):
        b, r = ctypes.c_uint64(), ctypes.c_uint64()
        ptr = d_out.ptr if isinstance(d_out, DeviceBuffer) else d_out
        self._ck(self.lib.sidgpu_emit_csv(self.h, site_begin, n_sites, ptr, out_cap, ctypes.byref(b), ctypes.byref(r)))
        return b.value, r.value

    def emit_columns(self, site_begin, n_sites, keep_dropped=False, strands=False):
        """sidgpu_emit_columns: the rows of the session as columns (numpy arrays) instead of CSV text.
        'chrom' is dictionary encoded: chrom_codes index into chrom_names.  Sites the method drops
        (coverage < 4 for bayes / likelihood_ratio, label 255) are filtered unless keep_dropped.
        strands (sessions begun with make_params(..., strands=True)): also 'profile' and 'fwd', the site's counts and those
        of the forward strand as (n, 4) arrays A, C, G, T."""
        n = int(n_sites)
        bufs = {"pos": DeviceBuffer(self, max(4 * n, 4)), "name_ref": DeviceBuffer(self, max(4 * n, 4)), "label": DeviceBuffer(self, max(n, 1)),
                "gt": DeviceBuffer(self, max(2 * n, 2)), "hom": DeviceBuffer(self, max(8 * n, 8)), "het": DeviceBuffer(self, max(8 * n, 8))}
        if strands:
            bufs["profile"] = DeviceBuffer(self, max(8 * n, 8))
            bufs["fwd"] = DeviceBuffer(self, max(8 * n, 8))
        try:
            cols = Columns(bufs["pos"].ptr, bufs["name_ref"].ptr, bufs["label"].ptr, bufs["gt"].ptr, bufs["hom"].ptr, bufs["het"].ptr,
                           bufs["profile"].ptr if strands else None, bufs["fwd"].ptr if strands else None)
            self._ck(self.lib.sidgpu_emit_columns(self.h, site_begin, n, ctypes.byref(cols)))
            if strands:
                unpack = lambda a: np.stack([(a >> np.uint64(16 * i)) & np.uint64(0xFFFF) for i in range(4)], axis=1).astype(np.uint16)
                profile4 = unpack(bufs["profile"].download(np.uint64, n))
                fwd4 = unpack(bufs["fwd"].download(np.uint64, n))
            pos = bufs["pos"].download(np.int32, n)
            refs = bufs["name_ref"].download(np.uint32, n)
            label = bufs["label"].download(np.uint8, n)
            gt = bufs["gt"].download(np.uint8, 2 * n).reshape(-1, 2)
            hom = bufs["hom"].download(np.float64, n)
            het = bufs["het"].download(np.float64, n)
        finally:
            for b in bufs.values():
                b.free()
        d_names, nbytes = ctypes.c_void_p(), ctypes.c_uint64()
        self._ck(self.lib.sidgpu_names(self.h, ctypes.byref(d_names), ctypes.byref(nbytes)))
        pool = self._download(d_names.value, np.uint8, nbytes.value).tobytes()
        uniq, codes = np.unique(refs, return_inverse=True)
        names = []
        for r in uniq:
            r = int(r)
            ln = pool[r] | (pool[r + 1] << 8)
            names.append(pool[r + 2:r + 2 + ln].decode("latin-1"))
        out = {"chrom_codes": codes.astype(np.int32), "chrom_names": names, "pos": pos, "label": label, "gt": gt, "hom_conf": hom, "het_conf": het}
        if strands:
            out["profile"], out["fwd"] = profile4, fwd4
        if not keep_dropped:
            keep = label != 255
            for k in ("chrom_codes", "pos", "label", "gt", "hom_conf", "het_conf") + (("profile", "fwd") if strands else ()):
                out[k] = out[k][keep]
        return out

    def emit_records(self, site_begin, n_sites):
        lab = DeviceBuffer(self, max(n_sites, 1))
        gt = DeviceBuffer(self, max(2 * n_sites, 2))
        hom = DeviceBuffer(self, max(8 * n_sites, 8))
        het = DeviceBuffer(self, max(8 * n_sites, 8))
        try:
            self._ck(self.lib.sidgpu_emit_records(self.h, site_begin, n_sites, lab.ptr, gt.ptr, hom.ptr, het.ptr))
            return (lab.download(np.uint8, n_sites), gt.download(np.uint8, 2 * n_sites).reshape(-1, 2),
                    hom.download(np.float64, n_sites), het.download(np.float64, n_sites))
        finally:
            for b in (lab, gt, hom, het):
                b.free()

    def set_global_histogram(self, d_profiles, d_counts, n):
        """The histograms of all shards (device pointers, n (profile, count) pairs, count 0 = padding): merged on the
        device; finish() then fits and classifies as one GPU holding the whole genome would."""
        self._ck(self.lib.sidgpu_set_global_histogram(self.h, d_profiles, d_counts, n))

    def histogram_device(self, min_coverage=4):
        """(n_unique, device pointer of the profiles, device pointer of the counts) of the session's histogram."""
        v = UniqueView()
        self._ck(self.lib.sidgpu_histogram(self.h, min_coverage, ctypes.byref(v)))
        return int(v.n_unique), v.d_profile, v.d_count

    def session_fit(self):
        f = Fit()
        nd = (ctypes.c_double * 4)()
        nu = ctypes.c_uint64()
        self._ck(self.lib.sidgpu_session_fit(self.h, ctypes.byref(f), nd, ctypes.byref(nu)))
        return {"pi": f.pi, "eps": f.eps, "fval": f.fval, "iterations": f.iterations, "evaluations": f.evaluations,
                "converged": bool(f.converged), "nd": list(nd), "n_unique": nu.value}

    # ---- K3 / K4 / K5
    def histogram(self, min_coverage=0):
        v = UniqueView()
        self._ck(self.lib.sidgpu_histogram(self.h, min_coverage, ctypes.byref(v)))
        return (self._download(v.d_profile, np.uint64, v.n_unique), self._download(v.d_count, np.uint64, v.n_unique),
                list(v.nd))

    def histogram_sums(self, min_coverage=4):
        """(n_unique, [A, C, G, T, total] integer sums) of the session's histogram: what ranks all-reduce."""
        v = UniqueView()
        self._ck(self.lib.sidgpu_histogram(self.h, min_coverage, ctypes.byref(v)))
        return v.n_unique, [int(x) for x in v.nd_sums]

    def count_unique(self, profiles, counts=None, min_coverage=0):
        """countUniqueProfiles (+ computeNucleotideDistribution) on packed profiles."""
        p = np.ascontiguousarray(profiles, dtype=np.uint64)
        v = UniqueView()
        if p.size == 0:
            return np.zeros(0, np.uint64), np.zeros(0, np.uint64), [0.25] * 4
        dp = DeviceBuffer(self, p.nbytes).upload(p)
        dc = None
        try:
            if counts is None:
                self._ck(self.lib.sidgpu_count_unique(self.h, dp.ptr, p.size, min_coverage, ctypes.byref(v)))
            else:
                c = np.ascontiguousarray(counts, dtype=np.uint64)
                dc = DeviceBuffer(self, c.nbytes).upload(c)
                self._ck(self.lib.sidgpu_count_unique_weighted(self.h, dp.ptr, dc.ptr, p.size, min_coverage, ctypes.byref(v)))
            return (self._download(v.d_profile, np.uint64, v.n_unique), self._download(v.d_count, np.uint64, v.n_unique), list(v.nd))
        finally:
            dp.free()
            if dc:
                dc.free()

    def set_fit(self, pi, eps, nd):
        a = (ctypes.c_double * 4)(*nd)
        self._ck(self.lib.sidgpu_set_fit(self.h, pi, eps, a))

    def finish_global(self, profiles_sorted):
        """sidgpu_finish for a sharded likelihood_ratio session: the merged unique profiles of all shards,
        in the reference's lexicographic order."""
        p = np.ascontiguousarray(profiles_sorted, dtype=np.uint64)
        self._ck(self.lib.sidgpu_finish_global(self.h, p.ctypes.data if p.size else None, p.size))

    def lynch_objective_partial(self, nd, pi, eps, d_out):
        """Leaves this rank's -log likelihood sum at device pointer d_out (no host sync)."""
        a = (ctypes.c_double * 4)(*nd)
        self._ck(self.lib.sidgpu_lynch_objective_partial(self.h, a, pi, eps, d_out))

    def lynch_objective(self, nd, pi, eps):
        a = (ctypes.c_double * 4)(*nd)
        out = ctypes.c_double()
        self._ck(self.lib.sidgpu_lynch_objective(self.h, a, pi, eps, ctypes.byref(out)))
        return out.value

    def lynch_fit(self, nd):
        a = (ctypes.c_double * 4)(*nd)
        f = Fit()
        self._ck(self.lib.sidgpu_lynch_fit(self.h, a, ctypes.byref(f)))
        return {"pi": f.pi, "eps": f.eps, "fval": f.fval, "iterations": f.iterations, "evaluations": f.evaluations,
                "converged": bool(f.converged)}

    def bh_adjust(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        if p.size == 0:
            return p.copy()
        d_in = DeviceBuffer(self, p.nbytes).upload(p)
        d_out = DeviceBuffer(self, p.nbytes)
        try:
            self._ck(self.lib.sidgpu_bh_adjust(self.h, d_in.ptr, p.size, d_out.ptr))
            return d_out.download(np.float64, p.size)
        finally:
            d_in.free()
            d_out.free()

    def format_g(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        if v.size == 0:
            return []
        d_in = DeviceBuffer(self, v.nbytes).upload(v)
        d_out = DeviceBuffer(self, 16 * v.size)
        try:
            self._ck(self.lib.sidgpu_format_g(self.h, d_in.ptr, v.size, d_out.ptr))
            raw = d_out.download(np.uint8, 16 * v.size).reshape(-1, 16)
            return [bytes(r).split(b"\0")[0].decode() for r in raw]
        finally:
            d_in.free()
            d_out.free()

    # ---- the one-call host path (what the `sid` binary does)
    def call_host(self, text, params, csv_capacity=None):
        """Returns (csv_rows_bytes, n_sites, n_rows); the header line is not included."""
        a = _as_u8(text)
        cap = int(csv_capacity) if csv_capacity else max(4096, a.nbytes + a.nbytes // 2)
        while True:
            out = np.empty(cap, dtype=np.uint8)
            nb, ns, nr = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
            rc = self.lib.sidgpu_call_host(self.h, ctypes.byref(params), a.ctypes.data if a.nbytes else None, a.nbytes,
                                           out.ctypes.data, cap, ctypes.byref(nb), ctypes.byref(ns), ctypes.byref(nr))
            if rc == 6 and nb.value > cap:      # SIDGPU_ECAPACITY: retry with the size the library reported
                cap = nb.value + 4096
                continue
            self._ck(rc)
            return out[:nb.value].tobytes(), ns.value, nr.value

    def call_host_bgzf(self, comp, params, csv_capacity=None):
        """call_host for the bytes of a BGZF file (its members are inflated on the device): (csv_rows_bytes, n_sites, n_rows)."""
        a = _as_u8(comp)
        cap = int(csv_capacity) if csv_capacity else max(4096, 6 * a.nbytes)
        while True:
            out = np.empty(cap, dtype=np.uint8)
            nb, ns, nr = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
            rc = self.lib.sidgpu_call_host_bgzf(self.h, ctypes.byref(params), a.ctypes.data if a.nbytes else None, a.nbytes,
                                                out.ctypes.data, cap, ctypes.byref(nb), ctypes.byref(ns), ctypes.byref(nr))
            if rc == 6 and nb.value > cap:
                cap = nb.value + 4096
                continue
            self._ck(rc)
            return out[:nb.value].tobytes(), ns.value, nr.value

    def bgzf_scan(self, comp, text_cap=1 << 62, max_blocks=None):
        """Member table of BGZF bytes (host only): (blocks array, consumed bytes, text bytes)."""
        buf = np.frombuffer(comp, dtype=np.uint8)
        cap = max_blocks or (len(comp) // 28 + 1)
        blocks = (_lib.BgzfBlock * cap)()
        n, used, tb = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
        self._ck(self.lib.sidgpu_bgzf_scan(buf.ctypes.data if len(comp) else None, len(comp), blocks, cap, text_cap,
                                           ctypes.byref(n), ctypes.byref(used), ctypes.byref(tb)))
        return blocks, n.value, used.value, tb.value

    def inflate_bgzf(self, comp):
        """BGZF bytes -> their text, inflated on the device (one warp per member)."""
        blocks, n, used, tb = self.bgzf_scan(comp)
        if used != len(comp):
            raise ValueError("truncated BGZF member")
        d_comp = DeviceBuffer(self, ((len(comp) + 15) & ~15) + 16)
        d_text = DeviceBuffer(self, tb + 16)
        try:
            if len(comp):
                d_comp.upload(np.frombuffer(comp, dtype=np.uint8))
            self._ck(self.lib.sidgpu_inflate_bgzf(self.h, d_comp.ptr, len(comp), blocks, n, d_text.ptr, tb))
            return d_text.download(np.uint8, tb).tobytes()
        finally:
            d_comp.free()
            d_text.free()

    def call_io(self, reader, writer, params, rewind=None, bgzf=False):
        """The streaming host path (sidgpu_call_io): `reader(n)` returns up to n bytes of pileup text (b"" at the
        end), `writer(rows)` receives CSV rows in file order, `rewind()` restarts the input (quality with -R).
        bgzf: the reader delivers the bytes of a BGZF file; its members are inflated on the device (sidgpu_call_io_bgzf).
        Returns (csv_bytes, n_sites, n_rows)."""
        failure = []

        def c_read(user, dst, cap):
            try:
                data = reader(min(cap, 1 << 24))
                if data:
                    ctypes.memmove(dst, data, len(data))
                return len(data)
            except Exception as e:          # an exception must not cross the C frame
                failure.append(e)
                return -1

        def c_write(user, rows, n):
            try:
                writer(ctypes.string_at(rows, n))
                return 0
            except Exception as e:
                failure.append(e)
                return 1

        def c_rewind(user):
            try:
                rewind()
                return 0
            except Exception as e:
                failure.append(e)
                return 1

        io = _lib.Io(_lib.IO_READ(c_read), _lib.IO_WRITE(c_write), _lib.IO_REWIND(c_rewind) if rewind else _lib.IO_REWIND(), None)
        nb, ns, nr = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        call = self.lib.sidgpu_call_io_bgzf if bgzf else self.lib.sidgpu_call_io
        rc = call(self.h, ctypes.byref(params), ctypes.byref(io), ctypes.byref(nb), ctypes.byref(ns), ctypes.byref(nr))
        if failure:
            raise failure[0]
        self._ck(rc)
        return nb.value, ns.value, nr.value


def parse_csv_rows(rows):
    """CSV rows (no header) -> list of OutputRecord (call.hpp:23-27)."""
    out = []
    for line in rows.split(b"\n"):
        if not line:
            continue
        f = line.decode("latin-1").rsplit(",", 6)
        out.append(OutputRecord(f[0], int(f[1]), f[2], f[3], float(f[4]), float(f[5]), f[6]))
    return out


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


def _call(text, params, ctx):
    ctx = ctx or default_context()
    rows, _, _ = ctx.call_host(text, params)
    return parse_csv_rows(rows)


# call.hpp:40-43 -- same names, same argument order and meaning.
def callSiteMLError(text, estimate_prior=False, prior=-1.0, error_threshold=0.1, significance_level=0.05, ctx=None):
    return _call(text, Context.make_params("local", estimate_prior, prior, error_threshold, significance_level), ctx)


def callBayes(text, ctx=None):
    return _call(text, Context.make_params("bayes"), ctx)


def callLikelihoodRatio(text, use_prior=False, significance_level=0.05, ctx=None):
    return _call(text, Context.make_params("likelihood_ratio", use_prior, -1.0, 0.1, significance_level), ctx)


def callQualityBasedSimple(text, estimate_prior=False, prior=-1.0, significance_level=0.05, ctx=None):
    return _call(text, Context.make_params("quality", estimate_prior, prior, 0.1, significance_level), ctx)


def sid_csv(text, method="local", estimate_prior=False, prior=-1.0, error_threshold=0.1, significance_level=0.05, ctx=None,
            het_only=False):
    """What `sid -m METHOD file` prints on stdout (sid.cpp:92-105), as bytes; with het_only what is left of it
    after the pipeline's `grep ',het,'` (scripts/sid-pipeline/run-sid.sh:16-17), header excluded."""
    if method not in METHODS:
        return CSV_HEADER                       # sid.cpp:92-100 has no else branch: header only
    ctx = ctx or default_context()
    rows, _, _ = ctx.call_host(text, Context.make_params(method, estimate_prior, prior, error_threshold, significance_level,
                                                         het_only=het_only))
    return rows if het_only else CSV_HEADER + rows


def call_columns(text, method="local", estimate_prior=False, prior=-1.0, error_threshold=0.1, significance_level=0.05, ctx=None, fit=None,
                 strands=False):
    """The rows `sid -m METHOD` prints, as columns: a dict of numpy arrays (see Context.emit_columns).
    All four methods (`quality` without -R: the text is fed once); the text is fed from device memory in one piece.
    strands: also the counts of every site and those of its forward strand ('profile', 'fwd')."""
    ctx = ctx or default_context()
    d = ctx.upload_text(text)
    try:
        ctx.begin(Context.make_params(method, estimate_prior, prior, error_threshold, significance_level, fit=fit, strands=strands))
        n = ctx.feed(d, d.text_len)
        if not ctx_streams(method, estimate_prior):
            ctx.finish()
        return ctx.emit_columns(0, n, strands=strands)
    finally:
        d.free()


def ctx_streams(method, estimate_prior):
    """Sessions that emit rows feed by feed (no genome-wide fit first)."""
    return method in ("local", "quality") and not estimate_prior


def columns_to_arrow(cols):
    """pyarrow.Table of the columns: chrom (dictionary), pos, label ('hom'/'het'), gt, hom_conf, het_conf."""
    import pyarrow as pa
    gt = [bytes(g).decode("latin-1") for g in cols["gt"]]
    chrom = pa.DictionaryArray.from_arrays(pa.array(cols["chrom_codes"], type=pa.int32()), pa.array(cols["chrom_names"], type=pa.string()))
    label = pa.DictionaryArray.from_arrays(pa.array(cols["label"].astype(np.int8)), pa.array(["hom", "het"]))
    table = {"chrom": chrom, "pos": pa.array(cols["pos"]), "label": label, "gt": pa.array(gt), "hom_conf": pa.array(cols["hom_conf"]),
             "het_conf": pa.array(cols["het_conf"])}
    if "fwd" in cols:            # strand-aware: counts per letter on the forward and on the reverse strand
        rev = (cols["profile"].astype(np.int32) - cols["fwd"].astype(np.int32)) & 0xFFFF
        for i, b in enumerate("ACGT"):
            table["fwd_" + b] = pa.array(cols["fwd"][:, i].astype(np.uint16))
            table["rev_" + b] = pa.array(rev[:, i].astype(np.uint16))
    return pa.table(table)
