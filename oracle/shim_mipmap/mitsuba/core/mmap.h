/* stand-in for <mitsuba/core/mmap.h>: a "memory-mapped file" over a buffer registered by the shim (mipmap_ref_shim.cpp) */
#pragma once
#include <mitsuba/mitsuba.h>
namespace mitsuba {
class MemoryMappedFile : public Object {
public:
    static void *&registeredData() { static void *p = NULL; return p; }
    static size_t &registeredSize() { static size_t s = 0; return s; }
    MemoryMappedFile(const fs::path &) : m_data(registeredData()), m_size(registeredSize()) {}
    MemoryMappedFile(const fs::path &, size_t size) : m_data(registeredData()), m_size(size) {}
    static ref<MemoryMappedFile> createTemporary(size_t) { return NULL; }
    void *getData() { return m_data; }
    size_t getSize() const { return m_size; }
private:
    void *m_data;
    size_t m_size;
};
}
