// fr constants in Montgomery form (generated with Python big integers; re-derived by the KAT tests).
#pragma once
#include <cstdint>
namespace b200zk {
// g = 19103219067921713944291392827692070036145651957329286315305642004821462161904 (order 2^28), gnark-crypto v0.9.1 fft.NewDomain root of unity
static __device__ __constant__ const uint32_t FR_ROOT28[8] = {0x80d13d9cu, 0x636e7355u, 0x2445ffd6u, 0xa22bf374u, 0x1eb203d8u, 0x56452ac0u, 0x2963f9e7u, 0x1860ef94u};
static __device__ __constant__ const uint32_t FR_ROOT28_INV[8] = {0x584bb683u, 0x89bcc016u, 0x0164a50cu, 0xe8d9887fu, 0x795eda3du, 0x755e95cbu, 0x1323b130u, 0x0f572b87u};
// fft.Domain.FrMultiplicativeGen = 5 and its inverse
static __device__ __constant__ const uint32_t FR_COSET[8] = {0x9fffffe6u, 0x1b0d0ef9u, 0xa32a913fu, 0xeaba68a3u, 0xd8dd0689u, 0x47d8eb76u, 0x20f5bbc3u, 0x15d00855u};
static __device__ __constant__ const uint32_t FR_COSET_INV[8] = {0x09999999u, 0xd7453974u, 0x83c3efa8u, 0xb4ada7d4u, 0xe57f3161u, 0xc49ca2f8u, 0xac156cb3u, 0x162a3754u};
// 1/2
static __device__ __constant__ const uint32_t FR_INV2[8] = {0x1ffffffeu, 0x783c14d8u, 0x0c8d1eddu, 0xaf982f6fu, 0xfcfd4f45u, 0x8f5f7492u, 0x3d9cbfacu, 0x1f37631au};
}  // namespace b200zk
