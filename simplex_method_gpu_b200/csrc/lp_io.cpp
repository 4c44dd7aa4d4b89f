// LP file input/output (include/b200lp_io.h).  Host code only.
// Text parser: the reference reads one token at a time with operator>> (load_matrix_impl,
// src/v4_cub_reduction.cu:94-104; header at v4:401-405).  Here the file is read once, cut into
// per-thread pieces at whitespace, tokens are counted per piece, and every piece is parsed with
// std::from_chars straight into its place of the column-major arrays.
#include "../../include/b200lp.h"
#include "../../include/b200lp_io.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

extern "C" int b200lp_internal_fail(int code, const char* msg);   // engine.cu: sets b200lp_last_error()

namespace {

const char MAGIC[8] = {'B', '2', '0', '0', 'L', 'P', '1', '\0'};

struct Header {
	char magic[8];
	int32_t dtype;
	int32_t reserved;
	int64_t m, n;
	char pad[32];
};
static_assert(sizeof(Header) == 64, "binary header is 64 bytes");

inline bool is_space(char ch) { return ch == ' ' || ch == '\n' || ch == '\t' || ch == '\r' || ch == '\f' || ch == '\v'; }

size_t elem_size(int32_t dtype) { return dtype == B200LP_F64 ? 8 : 4; }

void* alloc_host(size_t bytes, bool pinned, bool* was_pinned) {
	*was_pinned = false;
	if (bytes == 0) bytes = 1;
	if (pinned) {
		void* p = nullptr;
		if (cudaMallocHost(&p, bytes) == cudaSuccess) { *was_pinned = true; return p; }
		cudaGetLastError();
	}
	return std::malloc(bytes);
}

// bit 0 of `reserved` remembers how A was allocated
void free_host(void* p, bool pinned) {
	if (!p) return;
	if (pinned) cudaFreeHost(p); else std::free(p);
}

unsigned thread_count(size_t work_items) {
	unsigned hw = std::max(1u, std::thread::hardware_concurrency());
	return (unsigned)std::max<size_t>(1, std::min<size_t>(hw, work_items / (1 << 16) + 1));
}

template <typename F>
void parallel_for(unsigned nt, F&& fn) {
	if (nt <= 1) { fn(0u); return; }
	std::vector<std::thread> th;
	for (unsigned t = 0; t < nt; ++t) th.emplace_back(fn, t);
	for (auto& t : th) t.join();
}

// token k of the number stream (after the header) -> where it goes
template <typename T>
struct Sink {
	T *A, *b, *c;
	int64_t m, n;
	inline void put(int64_t k, T v) const {
		const int64_t mn = m * n;
		if (k < mn) A[(size_t)(k / n) + (size_t)(k % n) * (size_t)m] = v;     // row-major text -> column-major (v4:98)
		else if (k < mn + m) b[k - mn] = v;
		else c[k - mn - m] = v;
	}
};

std::string failed_at(int64_t k, int64_t m, int64_t n) {
	const int64_t mn = m * n;
	char buf[96];
	if (k < mn) std::snprintf(buf, sizeof buf, "Failed to read (%lld,%lld) for A", (long long)(k / n), (long long)(k % n));
	else if (k < mn + m) std::snprintf(buf, sizeof buf, "Failed to read (%lld,0) for b", (long long)(k - mn));
	else std::snprintf(buf, sizeof buf, "Failed to read (0,%lld) for c", (long long)(k - mn - m));
	return buf;
}

template <typename T>
int parse_text(const char* txt, size_t len, int32_t dtype, bool pinned, b200lp_problem* out) {
	// header: two integers (v4:401-405)
	const char* p = txt;
	const char* end = txt + len;
	long long m = 0, n = 0;
	auto read_int = [&](long long& v) {
		while (p < end && is_space(*p)) ++p;
		auto r = std::from_chars(p, end, v);
		if (r.ec != std::errc() || r.ptr == p) return false;
		p = r.ptr;
		return true;
	};
	if (!read_int(m) || !read_int(n) || m > n || m <= 0)
		return b200lp_internal_fail(B200LP_ERR_ARG, "Either failed to read m and n, or m > n.");
	const int64_t need = m * n + m + n;

	bool pin = false;
	T* A = (T*)alloc_host((size_t)m * n * sizeof(T), pinned, &pin);
	T* b = (T*)std::malloc((size_t)m * sizeof(T));
	T* c = (T*)std::malloc((size_t)n * sizeof(T));
	if (!A || !b || !c) { free_host(A, pin); std::free(b); std::free(c); return b200lp_internal_fail(B200LP_ERR_ARG, "out of host memory"); }
	Sink<T> sink{A, b, c, m, n};

	// cut [p, end) into pieces at whitespace
	const size_t body = (size_t)(end - p);
	const unsigned nt = thread_count(body);
	std::vector<const char*> cut(nt + 1);
	cut[0] = p;
	cut[nt] = end;
	for (unsigned t = 1; t < nt; ++t) {
		const char* q = p + body * t / nt;
		while (q < end && !is_space(*q)) ++q;     // never split a token
		cut[t] = q;
	}
	// pass 1: tokens per piece
	std::vector<int64_t> ntok(nt + 1, 0);
	parallel_for(nt, [&](unsigned t) {
		int64_t k = 0;
		bool in = false;
		for (const char* q = cut[t]; q < cut[t + 1]; ++q) {
			const bool sp = is_space(*q);
			k += (!sp && !in);
			in = !sp;
		}
		ntok[t + 1] = k;
	});
	for (unsigned t = 0; t < nt; ++t) ntok[t + 1] += ntok[t];
	// pass 2: parse; the first token that is not a number ends the stream like a failed operator>> would
	std::vector<int64_t> bad(nt, -1);
	parallel_for(nt, [&](unsigned t) {
		int64_t k = ntok[t];
		const char* q = cut[t];
		const char* e = cut[t + 1];
		while (k < need) {
			while (q < e && is_space(*q)) ++q;
			if (q >= e) break;
			if (*q == '+') ++q;                    // operator>> accepts a leading '+', from_chars does not
			// parsed in the problem's own scalar type: the decimal is rounded ONCE, like the reference's
			// operator>>(real) (v4:94-104); going through double first can be 1 ulp off in float
			T v;
			auto r = std::from_chars(q, e, v);
			if (r.ptr == q || r.ec == std::errc::invalid_argument) { bad[t] = k; break; }
			if (r.ec == std::errc::result_out_of_range) {
				const std::string tok(q, r.ptr);
				v = sizeof(T) == 8 ? (T)std::strtod(tok.c_str(), nullptr) : (T)std::strtof(tok.c_str(), nullptr);
			}
			sink.put(k, v);
			++k;
			q = r.ptr;
			// "1.5abc": operator>> takes the 1.5 and fails on the next extraction; a token that is not
			// consumed to its end therefore fails at the FOLLOWING index (and never shifts later indices)
			if (q < e && !is_space(*q)) { bad[t] = k; break; }
		}
	});
	int64_t got = std::min<int64_t>(ntok[nt], need);
	for (unsigned t = 0; t < nt; ++t)
		if (bad[t] >= 0) { got = std::min(got, bad[t]); break; }
	if (got < need) {
		free_host(A, pin); std::free(b); std::free(c);
		return b200lp_internal_fail(B200LP_ERR_ARG, failed_at(got, m, n).c_str());
	}
	out->dtype = dtype;
	out->reserved = pin ? 1 : 0;
	out->m = m;
	out->n = n;
	out->A = A;
	out->b = b;
	out->c = c;
	return B200LP_OK;
}

template <typename D, typename S>
void convert(D* dst, const S* src, size_t count) {
	for (size_t i = 0; i < count; ++i) dst[i] = (D)src[i];
}

int read_binary(FILE* f, int32_t dtype, bool pinned, b200lp_problem* out) {
	Header h;
	if (std::fread(&h, sizeof h, 1, f) != 1 || std::memcmp(h.magic, MAGIC, 8) != 0 || h.m <= 0 || h.m > h.n ||
			(h.dtype != B200LP_F32 && h.dtype != B200LP_F64))
		return b200lp_internal_fail(B200LP_ERR_ARG, "Either failed to read m and n, or m > n.");
	const size_t m = (size_t)h.m, n = (size_t)h.n, es = elem_size(dtype), fs = elem_size(h.dtype);
	bool pin = false;
	void* A = alloc_host(m * n * es, pinned, &pin);
	void* b = std::malloc(m * es);
	void* c = std::malloc(n * es);
	if (!A || !b || !c) { free_host(A, pin); std::free(b); std::free(c); return b200lp_internal_fail(B200LP_ERR_ARG, "out of host memory"); }
	bool ok = true;
	const char* names[3] = {"A", "b", "c"};
	void* dst[3] = {A, b, c};
	const size_t cnt[3] = {m * n, m, n};
	int which = 0;
	for (; which < 3 && ok; ++which) {
		if (fs == es) {
			ok = std::fread(dst[which], es, cnt[which], f) == cnt[which];
		} else {
			std::vector<unsigned char> tmp(std::min<size_t>(cnt[which], (size_t)1 << 22) * fs);
			size_t done = 0;
			while (ok && done < cnt[which]) {
				const size_t k = std::min<size_t>(cnt[which] - done, tmp.size() / fs);
				ok = std::fread(tmp.data(), fs, k, f) == k;
				if (!ok) break;
				if (es == 8) convert((double*)dst[which] + done, (const float*)tmp.data(), k);
				else convert((float*)dst[which] + done, (const double*)tmp.data(), k);
				done += k;
			}
		}
		if (!ok) break;
	}
	if (!ok) {
		free_host(A, pin); std::free(b); std::free(c);
		return b200lp_internal_fail(B200LP_ERR_ARG, (std::string("Failed to read (0,0) for ") + names[which]).c_str());
	}
	out->dtype = dtype;
	out->reserved = pin ? 1 : 0;
	out->m = h.m;
	out->n = h.n;
	out->A = A;
	out->b = b;
	out->c = c;
	return B200LP_OK;
}

template <typename T>
void append_number(std::string& s, T v) {
	char buf[40];
	auto r = std::to_chars(buf, buf + sizeof buf, v);      // shortest representation that round-trips
	s.append(buf, r.ptr);
}

template <typename T>
int write_text(FILE* f, const b200lp_problem* p) {
	const T* A = (const T*)p->A;
	const T* b = (const T*)p->b;
	const T* c = (const T*)p->c;
	const size_t m = (size_t)p->m, n = (size_t)p->n;
	std::fprintf(f, "%lld %lld\n", (long long)p->m, (long long)p->n);
	// rows are formatted in parallel, batches of rows written in order
	const size_t batch = std::max<size_t>(1, std::min<size_t>(m, ((size_t)1 << 22) / std::max<size_t>(n, 1) + 1));
	for (size_t i0 = 0; i0 < m; i0 += batch) {
		const size_t i1 = std::min(m, i0 + batch);
		std::vector<std::string> lines(i1 - i0);
		const unsigned nt = (unsigned)std::min<size_t>(thread_count((i1 - i0) * n), i1 - i0);
		parallel_for(nt, [&](unsigned t) {
			for (size_t i = i0 + t; i < i1; i += nt) {
				std::string& s = lines[i - i0];
				s.reserve(n * 12);
				for (size_t j = 0; j < n; ++j) {
					if (j) s.push_back(' ');
					append_number(s, A[i + j * m]);
				}
				s.push_back('\n');
			}
		});
		for (auto& s : lines)
			if (std::fwrite(s.data(), 1, s.size(), f) != s.size()) return B200LP_ERR_ARG;
	}
	std::string s;
	for (size_t i = 0; i < m; ++i) { if (i) s.push_back(' '); append_number(s, b[i]); }
	s.push_back('\n');
	for (size_t j = 0; j < n; ++j) { if (j) s.push_back(' '); append_number(s, c[j]); }
	s.push_back('\n');
	return std::fwrite(s.data(), 1, s.size(), f) == s.size() ? B200LP_OK : B200LP_ERR_ARG;
}

bool valid(const b200lp_problem* p) {
	return p && p->A && p->b && p->c && p->m > 0 && p->m <= p->n && (p->dtype == B200LP_F32 || p->dtype == B200LP_F64);
}

} // namespace

extern "C" {

int b200lp_read_lp(const char* path, int32_t dtype, int32_t pinned, b200lp_problem* out) {
	if (!path || !out || (dtype != B200LP_F32 && dtype != B200LP_F64)) return b200lp_internal_fail(B200LP_ERR_ARG, "read_lp: bad arguments");
	std::memset(out, 0, sizeof *out);
	FILE* f = std::fopen(path, "rb");
	if (!f) return b200lp_internal_fail(B200LP_ERR_ARG, (std::string("Could not open ") + path + ".").c_str());   // v4:397
	char magic[8] = {0};
	const size_t got = std::fread(magic, 1, 8, f);
	std::rewind(f);
	int rc;
	if (got == 8 && std::memcmp(magic, MAGIC, 8) == 0) {
		rc = read_binary(f, dtype, pinned != 0, out);
	} else {
		std::fseek(f, 0, SEEK_END);
		const long long sz = std::ftell(f);
		std::rewind(f);
		std::vector<char> buf((size_t)std::max<long long>(sz, 0) + 1);
		const size_t len = std::fread(buf.data(), 1, (size_t)std::max<long long>(sz, 0), f);
		buf[len] = 0;
		rc = dtype == B200LP_F64 ? parse_text<double>(buf.data(), len, dtype, pinned != 0, out)
		                         : parse_text<float>(buf.data(), len, dtype, pinned != 0, out);
	}
	std::fclose(f);
	return rc;
}

int b200lp_write_lp_text(const char* path, const b200lp_problem* p) {
	if (!path || !valid(p)) return b200lp_internal_fail(B200LP_ERR_ARG, "write_lp_text: bad arguments");
	FILE* f = std::fopen(path, "wb");
	if (!f) return b200lp_internal_fail(B200LP_ERR_ARG, (std::string("Could not open ") + path + ".").c_str());
	const int rc = p->dtype == B200LP_F64 ? write_text<double>(f, p) : write_text<float>(f, p);
	const bool closed = std::fclose(f) == 0;
	if (rc != B200LP_OK || !closed) return b200lp_internal_fail(B200LP_ERR_ARG, "write_lp_text: short write");
	return B200LP_OK;
}

int b200lp_write_lp_binary(const char* path, const b200lp_problem* p) {
	if (!path || !valid(p)) return b200lp_internal_fail(B200LP_ERR_ARG, "write_lp_binary: bad arguments");
	FILE* f = std::fopen(path, "wb");
	if (!f) return b200lp_internal_fail(B200LP_ERR_ARG, (std::string("Could not open ") + path + ".").c_str());
	Header h;
	std::memset(&h, 0, sizeof h);
	std::memcpy(h.magic, MAGIC, 8);
	h.dtype = p->dtype;
	h.m = p->m;
	h.n = p->n;
	const size_t es = elem_size(p->dtype), m = (size_t)p->m, n = (size_t)p->n;
	bool ok = std::fwrite(&h, sizeof h, 1, f) == 1 && std::fwrite(p->A, es, m * n, f) == m * n &&
		std::fwrite(p->b, es, m, f) == m && std::fwrite(p->c, es, n, f) == n;
	ok = (std::fclose(f) == 0) && ok;
	return ok ? B200LP_OK : b200lp_internal_fail(B200LP_ERR_ARG, "write_lp_binary: short write");
}

void b200lp_free_problem(b200lp_problem* p) {
	if (!p) return;
	free_host(p->A, (p->reserved & 1) != 0);
	std::free(p->b);
	std::free(p->c);
	std::memset(p, 0, sizeof *p);
}

} // extern "C"
