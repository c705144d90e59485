// Host-side tests of the FleEncoder/FleDecoder drop-ins, following the reference's own
// fle-test.cc (ValidateFle :172-200, Fle.SpecificSequences :202-242, Fle.TestValues :246-262,
// Fle.TestRandomValues :264-275) and adding what the reference never tests: the
// encoded-domain predicates and the Get(skip)/Skip cursor.  Needs a B200 (no CPU path).
#include <stdint.h>
#include <stdlib.h>

#include <vector>

#include "host-test-util.h"
#include "util/fle-encoding.h"

using namespace impala;
using std::vector;

const int MAX_WIDTH = 32;

// fle-test.cc:172-200
void ValidateFle(const vector<int>& values, int bit_width, uint8_t* expected_encoding, int expected_len) {
  const int len = 64 * 1024;
  static uint8_t buffer[len];
  EXPECT_LE(expected_len, len);

  FleEncoder encoder(buffer, len, bit_width);
  for (size_t i = 0; i < values.size(); ++i) {
    bool result = encoder.Put(values[i]);
    EXPECT_TRUE(result);
  }
  int encoded_len = encoder.Flush();

  if (expected_len != -1) {
    EXPECT_EQ(encoded_len, expected_len);
  }
  if (expected_encoding != NULL) {
    EXPECT_TRUE(memcmp(buffer, expected_encoding, expected_len) == 0);
  }

  // Verify read
  FleDecoder decoder(buffer, len, bit_width);
  for (size_t i = 0; i < values.size(); ++i) {
    uint64_t val;
    bool result = decoder.Get(&val);
    EXPECT_TRUE(result);
    EXPECT_EQ(values[i], val);
  }
}

TEST(Fle, SpecificSequences) {
  const int len = 1024;
  uint8_t expected_buffer[len];
  vector<int> values;

  // Test 50 0' followed by 50 1's
  values.resize(100);
  for (int i = 0; i < 50; ++i) values[i] = 0;
  for (int i = 50; i < 100; ++i) values[i] = 1;

  // The byte images the reference has commented out (fle-test.cc:217-223) hold exactly here,
  // because unfilled rows are written as 0 instead of stale staging bytes.
  *(reinterpret_cast<int64_t*>(&expected_buffer[0])) = 0x3fff;
  *(reinterpret_cast<int64_t*>(&expected_buffer[8])) = 0xfffffffff0000000;
  ValidateFle(values, 1, expected_buffer, 16);
  *(reinterpret_cast<int64_t*>(&expected_buffer[0])) = 0x3fff;
  *(reinterpret_cast<int64_t*>(&expected_buffer[8])) = 0x0;
  *(reinterpret_cast<int64_t*>(&expected_buffer[16])) = 0xfffffffff0000000;
  *(reinterpret_cast<int64_t*>(&expected_buffer[24])) = 0x0;
  ValidateFle(values, 2, expected_buffer, 32);

  for (int width = 2; width <= MAX_WIDTH; ++width) {
    ValidateFle(values, width, NULL, -1);
  }

  // Test 100 0's and 1's alternating
  for (int i = 0; i < 100; ++i) values[i] = i % 2;
  *(reinterpret_cast<int64_t*>(&expected_buffer[0])) = 0x5555555555555555;
  *(reinterpret_cast<int64_t*>(&expected_buffer[8])) = 0x5555555550000000;
  ValidateFle(values, 1, expected_buffer, 16);
  for (int width = 2; width <= MAX_WIDTH; ++width) {
    ValidateFle(values, width, NULL, -1);
  }
}

// fle-test.cc:246-254
void TestFleValues(int bit_width, int num_vals, int value = -1) {
  const uint64_t mod = 1LL << bit_width;
  vector<int> values;
  for (int v = 0; v < num_vals; ++v) {
    values.push_back((value != -1) ? value : (v % mod));
  }
  ValidateFle(values, bit_width, NULL, -1);
}

TEST(Fle, TestValues) {
  for (int width = 1; width <= MAX_WIDTH; ++width) {
    TestFleValues(width, 1);
    TestFleValues(width, 1024);
    TestFleValues(width, 1024, 0);
    TestFleValues(width, 1024, 1);
  }
}

TEST(Fle, TestRandomValues) {
  srand(20151018);  // the reference seeds with time(0); fixed here so failures reproduce
  for (int width = 1; width <= MAX_WIDTH; ++width) {
    const uint64_t mod = 1LL << width;
    vector<int> values;
    for (int v = 0; v < 1024; ++v) values.push_back(rand() % mod);
    ValidateFle(values, width, NULL, -1);
  }
}

TEST(Fle, EncoderBufferGuard) {
  // Put() returns false once the buffer has no room for a spare group (fle-encoding.h:251,8318)
  const int bw = 5;
  uint8_t buffer[3 * bw * 8];  // guard = 3*bw - bw words: groups 0 and 1 can be started
  FleEncoder enc(buffer, sizeof(buffer), bw);
  int accepted = 0;
  for (int i = 0; i < 1000; ++i) {
    if (!enc.Put(i % 32)) break;
    ++accepted;
  }
  EXPECT_EQ(accepted, 64 * 3);
  EXPECT_EQ(enc.Flush(), 3 * bw * 8);
  FleEncoder empty(buffer, sizeof(buffer), bw);
  EXPECT_EQ(empty.Flush(), bw * 8);  // Flush always packs one group (:9806-9811)
}

static bool ScalarCmp(int op, uint64_t v, uint64_t lit) {
  switch (op) {
    case 0: return v == lit;
    case 1: return v < lit;
    case 2: return v <= lit;
    case 3: return v > lit;
    case 4: return v >= lit;
  }
  return false;
}

static void RunPred(FleDecoder& d, int op, int64_t n, boost::dynamic_bitset<>& bs, uint64_t lit) {
  switch (op) {
    case 0: d.Eq(n, bs, lit); break;
    case 1: d.Lt(n, bs, lit); break;
    case 2: d.Le(n, bs, lit); break;
    case 3: d.Gt(n, bs, lit); break;
    case 4: d.Ge(n, bs, lit); break;
  }
}

TEST(Fle, PredicatesMatchScalarCompare) {
  srand(7);
  static uint8_t buffer[64 * 1024];
  for (int width = 1; width <= MAX_WIDTH; ++width) {
    const uint64_t mod = 1LL << width;
    const int n = 64 * 40 + 37;
    vector<uint64_t> values;
    FleEncoder enc(buffer, sizeof(buffer), width);
    for (int v = 0; v < n; ++v) {
      values.push_back(((uint64_t)rand() * 2654435761u + rand()) % mod);
      EXPECT_TRUE(enc.Put(values.back()));
    }
    int len = enc.Flush();
    const uint64_t lits[3] = {values[5], mod / 2, mod - 1};
    const int starts[4] = {0, 13, 64, 100};
    for (int op = 0; op < 5; ++op) {
      for (int li = 0; li < 3; ++li) {
        for (int si = 0; si < 4; ++si) {
          FleDecoder dec(buffer, len, width);
          for (int i = 0; i < starts[si]; ++i) { uint64_t x; EXPECT_TRUE(dec.Get(&x)); }
          boost::dynamic_bitset<> bs;
          RunPred(dec, op, n - starts[si], bs, lits[li]);
          EXPECT_EQ((int)bs.size(), n - starts[si]);
          int bad = 0;
          for (int i = 0; i < n - starts[si]; ++i) bad += bs[i] != ScalarCmp(op, values[starts[si] + i], lits[li]);
          EXPECT_EQ(bad, 0);
          // predicates do not advance the cursor
          uint64_t x;
          EXPECT_TRUE(dec.Get(&x));
          EXPECT_EQ(x, values[starts[si]]);
        }
      }
    }
    // In, and the scanner's window + Skip loop
    vector<uint64_t> inl;
    for (int i = 0; i < 9; ++i) inl.push_back(values[rand() % n]);
    FleDecoder dec(buffer, len, width);
    int done = 0, bad = 0;
    while (done < n) {
      int w = std::min(1024, n - done);
      boost::dynamic_bitset<> bs;
      dec.In(w, bs, inl);
      for (int i = 0; i < w; ++i) {
        bool want = false;
        for (size_t j = 0; j < inl.size(); ++j) want = want || values[done + i] == inl[j];
        bad += bs[i] != want;
      }
      EXPECT_TRUE(dec.Skip(w) || done + w >= n);
      done += w;
    }
    EXPECT_EQ(bad, 0);
  }
}

TEST(Fle, GetWithSkip) {
  srand(11);
  static uint8_t buffer[64 * 1024];
  for (int width = 1; width <= MAX_WIDTH; width += 3) {
    const uint64_t mod = 1LL << width;
    const int n = 64 * 50;
    vector<uint64_t> values;
    FleEncoder enc(buffer, sizeof(buffer), width);
    for (int v = 0; v < n; ++v) {
      values.push_back(rand() % mod);
      enc.Put(values.back());
    }
    int len = enc.Flush();
    FleDecoder dec(buffer, len, width);
    int pos = -1;
    while (true) {
      int skip = rand() % 150;
      if (pos + skip + 1 >= n) break;
      uint64_t x = 0;
      EXPECT_TRUE(dec.Get(&x, skip));
      pos += skip + 1;
      EXPECT_EQ(x, values[pos]);
    }
    uint64_t x;
    FleDecoder dec2(buffer, len, width);
    for (int i = 0; i < n; ++i) EXPECT_TRUE(dec2.Get(&x));
    EXPECT_TRUE(!dec2.Get(&x));  // buffer guard (:407)
  }
}

int main() { return RunAllTests(); }
