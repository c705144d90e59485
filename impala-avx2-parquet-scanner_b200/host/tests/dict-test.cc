// Host-side tests of DictEncoder/DictDecoder, following the reference's dict-test.cc
// (ValidateDict :32-62 and the integer/float cases of :64-157) and adding the literal ->
// code predicate translation (dict-encoding.h:461-541), which the reference never tests.
#include <stdint.h>
#include <stdlib.h>

#include <set>
#include <vector>

#include "host-test-util.h"
#include "util/dict-encoding.h"

using namespace impala;
using std::vector;

// dict-test.cc:32-62
template<typename T>
void ValidateDict(const vector<T>& values, int fixed_buffer_byte_size) {
  std::set<T> values_set(values.begin(), values.end());

  DictEncoder<T> encoder(fixed_buffer_byte_size);
  for (size_t i = 0; i < values.size(); ++i) encoder.Put(values[i]);
  EXPECT_EQ(encoder.num_entries(), (int)values_set.size());

  vector<uint8_t> dict_buffer(encoder.dict_encoded_size() + 8);
  encoder.WriteDict(&dict_buffer[0]);

  int data_buffer_len = encoder.EstimatedDataEncodedSize();
  vector<uint8_t> data_buffer(data_buffer_len);
  int data_len = encoder.WriteData(&data_buffer[0], data_buffer_len);
  EXPECT_TRUE(data_len > 0);
  encoder.ClearIndices();

  DictDecoder<T> decoder(&dict_buffer[0], encoder.dict_encoded_size(), fixed_buffer_byte_size);
  decoder.SetData(&data_buffer[0], data_len);
  for (size_t i = 0; i < values.size(); ++i) {
    T j;
    EXPECT_TRUE(decoder.GetValue(&j));
    EXPECT_TRUE(values[i] == j);
  }
  // sorted dictionary => order-preserving codes (dict-encoding.h:394-406)
  for (size_t i = 1; i < decoder.dict().size(); ++i) EXPECT_TRUE(decoder.dict()[i - 1] < decoder.dict()[i]);
}

template<typename T>
void IncrementValue(T* t) { ++(*t); }

template<typename T>
void TestNumbers(int max_value, int repeat, int value_byte_size) {
  vector<T> values;
  for (T val = 0; val < max_value; IncrementValue(&val)) {
    for (int i = 0; i < repeat; ++i) values.push_back(val);
  }
  ValidateDict(values, value_byte_size);
}

template<typename T>
void TestNumbers(int value_byte_size) {
  TestNumbers<T>(100, 1, value_byte_size);
  TestNumbers<T>(1, 100, value_byte_size);
  TestNumbers<T>(1, 1, value_byte_size);
  TestNumbers<T>(1, 2, value_byte_size);
}

TEST(DictTest, TestNumbers) {  // dict-test.cc:120-133
  TestNumbers<int8_t>(sizeof(int8_t));
  TestNumbers<int16_t>(sizeof(int16_t));
  TestNumbers<int32_t>(sizeof(int32_t));
  TestNumbers<int64_t>(sizeof(int64_t));
  TestNumbers<float>(sizeof(float));
  TestNumbers<double>(sizeof(double));
}

TEST(DictTest, Predicates) {
  srand(3);
  const int n = 64 * 100 + 21;
  vector<int32_t> values;
  for (int i = 0; i < n; ++i) values.push_back((rand() % 300) * 7 - 500);  // sparse domain
  DictEncoder<int32_t> encoder(4);
  for (int i = 0; i < n; ++i) encoder.Put(values[i]);
  vector<uint8_t> dict_buffer(encoder.dict_encoded_size());
  encoder.WriteDict(&dict_buffer[0]);
  vector<uint8_t> data(encoder.EstimatedDataEncodedSize());
  int data_len = encoder.WriteData(&data[0], data.size());
  EXPECT_TRUE(data_len > 0);
  const int32_t lits[] = {-10000, -500, -499, 0, 2, 7 * 150 - 500, 1593, 1594, 100000};
  for (size_t li = 0; li < sizeof(lits) / sizeof(lits[0]); ++li) {
    int32_t lit = lits[li];
    for (int op = 0; op < 5; ++op) {
      DictDecoder<int32_t> dec(&dict_buffer[0], dict_buffer.size(), 4);
      dec.SetData(&data[0], data_len);
      boost::dynamic_bitset<> bs;
      switch (op) {
        case 0: dec.Eq(n, bs, lit); break;
        case 1: dec.Lt(n, bs, lit); break;
        case 2: dec.Le(n, bs, lit); break;
        case 3: dec.Gt(n, bs, lit); break;
        case 4: dec.Ge(n, bs, lit); break;
      }
      EXPECT_EQ((int)bs.size(), n);
      int bad = 0;
      for (int i = 0; i < n; ++i) {
        bool want = op == 0 ? values[i] == lit : op == 1 ? values[i] < lit : op == 2 ? values[i] <= lit
                  : op == 3 ? values[i] > lit : values[i] >= lit;
        bad += bs[i] != want;
      }
      EXPECT_EQ(bad, 0);
    }
  }
  vector<int32_t> inl;
  inl.push_back(values[3]); inl.push_back(12345); inl.push_back(values[77]); inl.push_back(-501);
  DictDecoder<int32_t> dec(&dict_buffer[0], dict_buffer.size(), 4);
  dec.SetData(&data[0], data_len);
  boost::dynamic_bitset<> bs;
  dec.In(n, bs, inl);
  int bad = 0;
  for (int i = 0; i < n; ++i) bad += bs[i] != (values[i] == values[3] || values[i] == values[77]);
  EXPECT_EQ(bad, 0);
  vector<int32_t> none;
  none.push_back(1); none.push_back(3);
  boost::dynamic_bitset<> bs2;
  dec.In(n, bs2, none);
  EXPECT_EQ((int)bs2.size(), n);
  EXPECT_EQ((int)bs2.count(), 0);
}

int main() { return RunAllTests(); }
