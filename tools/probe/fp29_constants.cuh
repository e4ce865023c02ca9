// radix-2^29 constants of the fp29 experiment (generated once by the field29() helper that used to live in tools/gen_constants.py)
#pragma once
#include "ptx.cuh"
namespace b2k {
// BLS12-381 base field in radix 2^29 (14 limbs, R = 2^406): MSM bucket arithmetic
struct Bls381Fp29 {
  static constexpr int N = 14;
  static constexpr int W = 29;
  static constexpr uint32_t PINV = 0x1ffcfffdu;   // -p^-1 mod 2^29
  B2K_D static uint32_t mod(int j) { constexpr uint32_t t[14] = {0x1fffaaabu, 0x0ff7ffffu, 0x14ffffeeu, 0x17fffd62u, 0x0f6241eau, 0x09507b58u, 0x0afd9cc3u, 0x109e70a2u, 0x1764774bu, 0x121a5d66u, 0x12c6e9edu, 0x12ffcd34u, 0x00111ea3u, 0x0000000du}; return t[j]; }
  B2K_D static uint32_t r1(int j) { constexpr uint32_t t[14] = {0x03a9fb84u, 0x0ba00690u, 0x071288f1u, 0x0f59bcc5u, 0x126cb614u, 0x0585bf36u, 0x1b85ac3du, 0x1cf856fau, 0x1891ecbdu, 0x1a7eec05u, 0x155a88f0u, 0x0741ac6du, 0x1317c30fu, 0x00000009u}; return t[j]; }
  B2K_D static uint32_t r2(int j) { constexpr uint32_t t[14] = {0x15bef7aeu, 0x1031cd0eu, 0x02dd93e8u, 0x09226323u, 0x0e6e2cd2u, 0x11684daau, 0x1170e5dbu, 0x088e25b1u, 0x1b366399u, 0x1c536f47u, 0x0d1f9cbcu, 0x0278b67fu, 0x1ea66a2bu, 0x0000000cu}; return t[j]; }
  B2K_D static uint32_t from_r384(int j) { constexpr uint32_t t[14] = {0x1fddebbdu, 0x1a4f5474u, 0x0291f399u, 0x14d03b3cu, 0x0f6cad2cu, 0x1b4cabcau, 0x1592827cu, 0x021c6ac7u, 0x1ec52a84u, 0x16fd5ec4u, 0x0c960da6u, 0x0fd2af6bu, 0x13263591u, 0x0000000bu}; return t[j]; }
  B2K_D static uint32_t to_r384(int j) { constexpr uint32_t t[14] = {0x0002fffdu, 0x10480000u, 0x0300009du, 0x08001788u, 0x158baebfu, 0x0c2ba9e3u, 0x1d157d22u, 0x0a6e0a4au, 0x0d77ce58u, 0x1d12b763u, 0x1701c6a5u, 0x1501c926u, 0x1f65ec3fu, 0x0000000au}; return t[j]; }
  B2K_D static uint32_t sub_c(int j) { constexpr uint32_t t[14] = {0x3ff55560u, 0x3efffffeu, 0x3ffffdceu, 0x3fffac53u, 0x2c483d56u, 0x2a0f6b0eu, 0x3fb39868u, 0x33ce1449u, 0x2c8ee96fu, 0x234bacd6u, 0x38dd3db1u, 0x3ff9a691u, 0x2223d471u, 0x0000019fu}; return t[j]; }
  B2K_D static uint32_t beta(int j) { constexpr uint32_t t[14] = {0x0abf79e7u, 0x194b14ebu, 0x1ceb16a6u, 0x1845cd5du, 0x0c814568u, 0x199ca109u, 0x13ee6ee1u, 0x05d123e5u, 0x0dfbdce2u, 0x010ecb54u, 0x0d9337edu, 0x1daaf4b0u, 0x146806fbu, 0x0000000cu}; return t[j]; }
};

}  // namespace b2k
