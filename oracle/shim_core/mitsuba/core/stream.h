/* Stand-in header (test infrastructure only, see oracle/shim_core/README): scaffolding that lets the reference's own sources and
 * headers compile where they lie under /root/reference, without boost or the rest of libcore.  No algorithm lives here. */
#pragma once
namespace mitsuba { class Stream { public: float readSingle(); double readDouble(); void writeSingle(float); void writeDouble(double); int readInt(); void writeInt(int); float readFloat(); void writeFloat(float);
 template <typename T> void readArray(T *, size_t); template <typename T> void writeArray(const T *, size_t); template <typename T> T readElement(); template <typename T> void writeElement(T);
 void readFloatArray(float *, size_t); void readULongArray(uint64_t *, size_t); void writeULongArray(const uint64_t *, size_t); void writeFloatArray(const float *, size_t); unsigned int readUInt(); void writeUInt(unsigned int); size_t readSize(); void writeSize(size_t); bool readBool(); void writeBool(bool); short readShort(); void writeShort(short); long long readLong(); void writeLong(long long); unsigned long long readULong(); void writeULong(unsigned long long); }; }
