// swift-tools-version:5.9
// Swift package that lets the reference's own XCTest files run against libairwave_cuda.so on a host that has a Swift
// toolchain and a B200 (this repository's build image has neither Swift nor macOS; see airwave_b200/swift/README.md).
//
// Layout expected by this manifest (created by airwave_b200/swift/assemble_package.sh):
//   Package.swift
//   Sources/CAirwaveCUDA/{module.modulemap, airwave_cuda.h}     <- include/ of this repository
//   Sources/AirwaveCUDA/AirwaveCUDA.swift                        <- airwave_b200/swift/AirwaveCUDA.swift
//   Tests/AirwaveCUDATests/{ConvolutionEngineTests.swift, RealtimeAudioProcessorTests.swift}   <- copied UNCHANGED from the
//                                                                   reference checkout (AirwaveTests/), `@testable import Airwave`
//                                                                   rewritten to `@testable import AirwaveCUDA` by the script
import PackageDescription

let libDir = Context.environment["AIRWAVE_CUDA_LIB_DIR"] ?? "../lib"

let package = Package(
    name: "AirwaveCUDA",
    products: [.library(name: "AirwaveCUDA", targets: ["AirwaveCUDA"])],
    targets: [
        .systemLibrary(name: "CAirwaveCUDA", path: "Sources/CAirwaveCUDA"),
        .target(
            name: "AirwaveCUDA",
            dependencies: ["CAirwaveCUDA"],
            path: "Sources/AirwaveCUDA",
            linkerSettings: [.unsafeFlags(["-L\(libDir)", "-lairwave_cuda", "-Xlinker", "-rpath", "-Xlinker", libDir])]
        ),
        .testTarget(name: "AirwaveCUDATests", dependencies: ["AirwaveCUDA"], path: "Tests/AirwaveCUDATests"),
    ]
)
